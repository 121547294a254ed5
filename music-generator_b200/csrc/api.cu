// C-ABI plumbing shared by all translation units: version + thread-local error text.
#include <stdarg.h>

#include "dj_common.cuh"

static thread_local char g_dj_error[512] = "";

void dj_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_dj_error, sizeof(g_dj_error), fmt, ap);
  va_end(ap);
}

// ---- deterministic reductions -----------------------------------------------------
// Kernels whose result is a sum over CTAs (split weight-gradient GEMMs, bias gradients, the conv gradient, column
// sums) either add with fp32 atomics -- order, hence rounding, varies from run to run -- or, when the caller has
// registered a workspace for the launching stream, write one partial per CTA and have dj_ordered_reduce add the
// partials in index order: the same bits every run.
namespace {
struct DjWs { void* stream; float* ptr; int64_t floats; };
thread_local DjWs g_ws[8] = {};

__global__ void ordered_reduce_kernel(const float* __restrict__ part, int P, int64_t pstride, int64_t rows, int cols,
                                      int64_t ldp, float* __restrict__ out, int64_t ldo) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    const float* src = part + r * ldp + c;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += src[(int64_t)p * pstride];     // fixed order
    out[r * ldo + c] += s;
  }
}
}  // namespace

extern "C" int dj_set_reduce_workspace(void* stream, float* ws, int64_t floats) {
  DJ_CHECK_ARG(floats >= 0 && (ws != nullptr || floats == 0), "dj_set_reduce_workspace: bad arguments");
  int slot = -1;
  for (int i = 0; i < 8; ++i)
    if (g_ws[i].floats > 0 && g_ws[i].stream == stream) slot = i;
  if (slot < 0)
    for (int i = 0; i < 8 && slot < 0; ++i)
      if (g_ws[i].floats == 0) slot = i;
  DJ_CHECK_ARG(slot >= 0, "dj_set_reduce_workspace: more than 8 streams registered");
  g_ws[slot] = DjWs{stream, ws, floats};
  return 0;
}

int dj_reduce_workspace(void* stream, int64_t need, float** out) {
  *out = nullptr;
  for (int i = 0; i < 8; ++i)
    if (g_ws[i].floats > 0 && g_ws[i].stream == stream) {
      // a registered workspace that is too small is an error, never a silent return to the atomics
      DJ_CHECK_ARG(g_ws[i].floats >= need, "deterministic reduction needs a workspace of %lld floats, %lld registered",
                   (long long)need, (long long)g_ws[i].floats);
      *out = g_ws[i].ptr;
      return 0;
    }
  return 0;
}

int dj_ordered_reduce(const float* part, int P, int64_t pstride, int64_t rows, int cols, int64_t ldp, float* out,
                      int64_t ldo, void* stream) {
  const int64_t total = rows * cols;
  int64_t blocks = (total + 255) / 256;
  if (blocks > dj_num_sms() * 8) blocks = dj_num_sms() * 8;
  ordered_reduce_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(part, P, pstride, rows, cols, ldp, out, ldo);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_version(void) { return 101; }
extern "C" const char* dj_last_error(void) { return g_dj_error; }
