"""numpy restatements of the small encoders and activations.  TEST INFRASTRUCTURE ONLY.

freq_encode   : freqencoder/src/freqencoder.cu:30-58 (channel order, scalbnf, sin(x + pi/2) for cos)
freq_backward : freqencoder/src/freqencoder.cu:63-94
sh_encode     : shencoder/src/shencoder.cu:49-70 (degree <= 4), same polynomials as testing/test_shencoder.py:51-92; bands 4..7
                (degrees 5..8, :71-135) from scipy's spherical harmonics, for unit vectors
trunc_exp     : activation.py:5-17
"""
import numpy as np


def freq_encode(x, degree):
    x = np.ascontiguousarray(x, np.float32)
    B, D = x.shape
    C = D + 2 * D * degree
    out = np.empty((B, C), np.float64)
    for c in range(C):
        if c < D:
            out[:, c] = x[:, c]
        else:
            col = c // D - 1
            d = c % D
            freq = col // 2
            phase = np.float32((col % 2) * (np.float32(3.141592653589793) / np.float32(2)))
            arg = (np.ldexp(x[:, d], freq).astype(np.float32) + phase).astype(np.float32)
            out[:, c] = np.sin(arg.astype(np.float64))
    return out


def freq_backward(grad, outputs, D, degree):
    grad = np.asarray(grad, np.float64)
    outputs = np.asarray(outputs, np.float64)
    B = grad.shape[0]
    gi = np.empty((B, D))
    for d in range(D):
        result = grad[:, d].copy()
        for f in range(degree):
            base = D + 2 * D * f
            result += (2.0 ** f) * (grad[:, base + d] * outputs[:, base + D + d] - grad[:, base + D + d] * outputs[:, base + d])
        gi[:, d] = result
    return gi


def sh_encode(dirs, degree=4):
    d = np.asarray(dirs, np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    out = np.empty((d.shape[0], degree * degree))
    out[:, 0] = 0.28209479177387814
    if degree > 1:
        out[:, 1] = -0.48860251190291987 * y
        out[:, 2] = 0.48860251190291987 * z
        out[:, 3] = -0.48860251190291987 * x
    if degree > 2:
        out[:, 4] = 1.0925484305920792 * xy
        out[:, 5] = -1.0925484305920792 * yz
        out[:, 6] = 0.94617469575755997 * z2 - 0.31539156525251999
        out[:, 7] = -1.0925484305920792 * xz
        out[:, 8] = 0.54627421529603959 * x2 - 0.54627421529603959 * y2
    if degree > 3:
        out[:, 9] = 0.59004358992664352 * y * (-3.0 * x2 + y2)
        out[:, 10] = 2.8906114426405538 * xy * z
        out[:, 11] = 0.45704579946446572 * y * (1.0 - 5.0 * z2)
        out[:, 12] = 0.3731763325901154 * z * (5.0 * z2 - 3.0)
        out[:, 13] = 0.45704579946446572 * x * (1.0 - 5.0 * z2)
        out[:, 14] = 1.4453057213202769 * z * (x2 - y2)
        out[:, 15] = 0.59004358992664352 * x * (-x2 + 3.0 * y2)
    if degree > 4:
        # bands 4..7 (shencoder.cu:71-135): not restated polynomial by polynomial — evaluated from scipy's complex spherical
        # harmonics (valid for UNIT vectors; pinned by the check that the same construction reproduces bands 0..3 above and, on the
        # GPU box, by the reference extension itself)
        out[:, 16:] = sh_bands_scipy(d, range(4, degree))
    return out


def sh_bands_scipy(dirs, bands):
    """Real spherical harmonics of UNIT vectors for the given bands, ordered l*l + l + m within a band, in the reference's
    convention: Y_l^{+m} = sqrt(2) Re(Y_l^m), Y_l^{-m} = sqrt(2) Im(Y_l^m) with scipy's Condon-Shortley-phased complex Y_l^m."""
    from scipy import special
    d = np.asarray(dirs, np.float64)
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
    theta = np.arccos(np.clip(d[:, 2], -1.0, 1.0))   # polar
    phi = np.arctan2(d[:, 1], d[:, 0])               # azimuth
    cols = []
    for l in bands:
        blk = np.empty((d.shape[0], 2 * l + 1))
        for m in range(0, l + 1):
            if hasattr(special, "sph_harm_y"):
                Y = special.sph_harm_y(l, m, theta, phi)
            else:
                Y = special.sph_harm(m, l, phi, theta)
            if m == 0:
                blk[:, l] = Y.real
            else:
                blk[:, l + m] = np.sqrt(2.0) * Y.real
                blk[:, l - m] = np.sqrt(2.0) * Y.imag
        cols.append(blk)
    return np.concatenate(cols, axis=1)


def sh_jacobian_fd(dirs, degree=4, eps=1e-6):
    """Central finite differences of sh_encode -> [B, 3, degree^2] (layout of dy_dx, shencoder.cu:139-141)."""
    d = np.asarray(dirs, np.float64)
    J = np.empty((d.shape[0], 3, degree * degree))
    for k in range(3):
        e = np.zeros(3)
        e[k] = eps
        J[:, k, :] = (sh_encode(d + e, degree) - sh_encode(d - e, degree)) / (2 * eps)
    return J


def trunc_exp_forward(x):
    return np.exp(np.asarray(x, np.float32).astype(np.float64))


def trunc_exp_backward(x, g):
    return np.asarray(g, np.float64) * np.exp(np.clip(np.asarray(x, np.float64), -15, 15))
