// api.cu — library identification and status strings of libseald_b200.so.
#include "common.cuh"

extern "C" int seald_version(void) { return 100; }

extern "C" int seald_sm_arch(void) {
#ifdef SEALD_SM_ARCH
    return SEALD_SM_ARCH;
#else
    return 100;
#endif
}

extern "C" const char* seald_strerror(int status) {
    switch (status) {
        case 0: return "ok";
        case SEALD_E_BADARG: return "bad argument (null pointer or inconsistent sizes)";
        case SEALD_E_UNSUPPORTED: return "unsupported configuration (D / C / width / degree / dtype)";
        case SEALD_E_ALIGN: return "pointer not aligned for the kernel's vector width";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown error";
}
