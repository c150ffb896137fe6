// api.cu -- library-wide entry points of libsis3d.so.
#include "common.cuh"
namespace sis3d { unsigned long long g_launch_count = 0; }

extern "C" const char *sis3d_strerror(int code) {
    switch (code) {
        case SIS3D_OK: return "ok";
        case SIS3D_EINVAL: return "invalid argument";
        case SIS3D_ELAUNCH: return "CUDA launch/runtime error";
        case SIS3D_EWORKSPACE: return "workspace too small";
        case SIS3D_EUNSUPPORTED: return "unsupported configuration";
        default: return "unknown sis3d error";
    }
}
extern "C" int sis3d_version(void) { return 100; }
extern "C" int64_t sis3d_launch_count(void) { return (int64_t)sis3d::g_launch_count; }
