// C-ABI: error slot + the stand-alone GEMM operator export (include/pi05.h).
#include <cstdio>
#include <cstring>

#include "../../include/pi05.h"
#include "errors.h"
#include "gemm.h"

namespace pi05 {
static thread_local char g_last_error[1024] = "";
void set_error(const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg ? msg : "");
}
const char* get_error() { return g_last_error; }
static unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
unsigned long long launch_count() { return g_launches; }
}  // namespace pi05

extern "C" {

int pi05_abi_version(void) { return PI05_ABI_VERSION; }
unsigned long long pi05_launch_count(void) { return pi05::launch_count(); }
void pi05_gemm_profile_enable(int on) { pi05::gemm_profile_enable(on); }
int pi05_gemm_profile_report(char* buf, int len) { return pi05::gemm_profile_report(buf, len); }
const char* pi05_last_error(void) { return pi05::get_error(); }

int pi05_gemm_bf16(const pi05_gemm_desc* d, void* stream) {
  if (!d) {
    pi05::set_error("pi05_gemm_bf16: null descriptor");
    return 1;
  }
  pi05::GemmArgs a;
  a.M = d->M; a.N = d->N; a.K = d->K; a.batch = d->batch > 0 ? d->batch : 1;
  a.batch_inner = d->batch_inner;
  a.a_batch_stride1 = d->a_batch_stride1; a.b_batch_stride1 = d->b_batch_stride1;
  a.d_batch_stride1 = d->d_batch_stride1; a.d2_batch_stride1 = d->d2_batch_stride1;
  a.res_batch_stride1 = d->res_batch_stride1;
  a.A = d->A; a.B = d->B; a.a_major = d->a_major; a.b_major = d->b_major;
  a.lda = d->lda; a.ldb = d->ldb; a.a_batch_stride = d->a_batch_stride; a.b_batch_stride = d->b_batch_stride;
  a.epilogue = d->epilogue; a.D = d->D; a.ldd = d->ldd; a.d_batch_stride = d->d_batch_stride;
  a.D2 = d->D2; a.ldd2 = d->ldd2; a.d2_batch_stride = d->d2_batch_stride;
  a.bias = d->bias; a.res = d->res; a.ldres = d->ldres; a.res_batch_stride = d->res_batch_stride;
  a.gate = d->gate; a.gate_rows = d->gate_rows; a.ldgate = d->ldgate;
  a.scale = d->scale; a.accumulate = d->accumulate; a.block_n = d->block_n;
  a.splitk_ws = static_cast<float*>(d->workspace); a.splitk_ws_bytes = d->workspace_bytes;
  char err[512] = "";
  int rc = pi05::gemm_bf16(a, static_cast<cudaStream_t>(stream), err, sizeof(err));
  if (rc != 0) pi05::set_error(err);
  return rc;
}

}  // extern "C"
