from .raymarching import (near_far_from_aabb, sph_from_ray, morton3D, morton3D_invert, packbits, march_rays_train,
                          composite_rays_train, march_rays, composite_rays, compact_alive, march_rays_seal, march_rays_train_seal, occupancy_aabb)

__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train", "march_rays", "composite_rays", "compact_alive", "march_rays_seal", "march_rays_train_seal", "occupancy_aabb"]
