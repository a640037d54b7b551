// api.cu -- library lifecycle, error reporting, global knobs.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dk {

static thread_local char g_err[512] = "";
static int g_sm_count = 148;
static int g_device = -1;
static int g_backend = 0;
static unsigned long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() { return g_sm_count; }
int gemm_backend() { return g_backend; }
void count_launch() { ++g_launches; }

int init_gemm_tcgen05();  // gemm_tcgen05.cu: opt-in shared memory sizes
int init_depthwise();     // depthwise.cu
int init_conv_rows();     // conv_rows.cu
int init_conv_tma();      // conv_tma.cu
int bn_fused_init();      // bn_fused.cu

}  // namespace dk

extern "C" {

int dk_version(void) { return 100; }  // 0.1.0

const char *dk_last_error(void) { return dk::g_err; }

int dk_init(int device) {
    DK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    DK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        dk::set_error("libdorknet_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return DK_ERR_UNSUPPORTED;
    }
    dk::g_sm_count = prop.multiProcessorCount;
    dk::g_device = device;
    int rc = dk::init_depthwise();
    if (rc) return rc;
    rc = dk::init_gemm_tcgen05();
    if (rc) return rc;
    rc = dk::init_conv_rows();
    if (rc) return rc;
    rc = dk::init_conv_tma();
    if (rc) return rc;
    rc = dk::bn_fused_init();
    if (rc) return rc;
    return DK_OK;
}

int dk_destroy(void) {
    dk::g_device = -1;
    return DK_OK;
}

int dk_sm_count(void) { return dk::g_sm_count; }

unsigned long long dk_kernel_launches(void) { return dk::g_launches; }

int dk_set_gemm_backend(int backend) {
    if (backend != 0 && backend != 1) {
        dk::set_error("dk_set_gemm_backend: backend must be 0 (tcgen05) or 1 (simt), got %d", backend);
        return DK_ERR_INVALID;
    }
    dk::g_backend = backend;
    return DK_OK;
}
int dk_get_gemm_backend(void) { return dk::g_backend; }

}  // extern "C"
