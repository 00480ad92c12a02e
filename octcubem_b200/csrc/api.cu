// Library-level entry points: version, thread-local error string, device query.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

static thread_local char g_err[512] = "";

void oct_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;  // kernels launched through this library (bench.py reports it)

int oct_check_launch(const char* what) {
  __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    oct_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return OCT_OK;
}

static int g_pdl_override = -1;  // oct_set_pdl(): -1 = follow the OCT_PDL environment variable

bool oct_pdl_enabled() {
  const int ov = __atomic_load_n(&g_pdl_override, __ATOMIC_RELAXED);
  if (ov >= 0) return ov != 0;
  static int on = -1;
  if (on < 0) { const char* e = getenv("OCT_PDL"); on = (e && e[0] == '1') ? 1 : 0; }  // opt-in: see common.cuh
  return on != 0;
}

extern "C" void oct_set_pdl(int mode) { __atomic_store_n(&g_pdl_override, mode < 0 ? -1 : (mode ? 1 : 0), __ATOMIC_RELAXED); }

int oct_num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

extern "C" uint64_t oct_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" const char* oct_version(void) { return "octcube_b200 0.1.0 (sm_100a)"; }
extern "C" const char* oct_last_error(void) { return g_err; }

extern "C" int oct_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { oct_set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { oct_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (p.major != 10) { oct_set_error("octcube_b200 needs an sm_100-class device, found sm_%d%d", p.major, p.minor); return OCT_ERR_UNSUPPORTED; }
  return OCT_OK;
}
