// Error plumbing shared by every translation unit of libsfattack.
#include <stdio.h>
#include <string.h>

#include "sfk_common.cuh"

static thread_local char g_err[256] = "ok";

void sfk_set_error(const char* msg) {
  strncpy(g_err, msg, sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

int sfk_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

extern "C" int sfk_version(void) { return 100; }
extern "C" const char* sfk_last_error_string(void) { return g_err; }

static int g_act_f32 = 0;
int sfk_act_f32() { return g_act_f32; }
extern "C" int sfk_set_activation_dtype(int f32) {
  g_act_f32 = f32 ? 1 : 0;
  return 0;
}
extern "C" int sfk_get_activation_dtype(void) { return g_act_f32; }
