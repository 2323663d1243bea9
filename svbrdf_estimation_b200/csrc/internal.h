// internal.h - launch helpers shared by kernels.cu and host_ctx.cu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

int svb_fail(int code, const char* msg);
int svb_cuda_status(cudaError_t e, const char* where);
int svb_check_shape(int B, int H, int W, int N);

// packed = two pixels per thread (W even, 8-byte aligned pointers); decides the CTA count per image
bool svb_loss_packed(int W, const void* input, const void* target, const void* grad, const void* lin);
int svb_ctas_per_image(int HW, bool packed);
int svb_launch_loss_range(const float* input, const float* target, float* grad, int B, int HW, int W,
                          const float* scenes, int N, const float* lin, float* part_render, float* part_l1,
                          bool mixed, float l1_weight, int b0, int bn, cudaStream_t st, bool packed, int lay = 0,
                          bool accurate = false);
int svb_layout_id(int input_layout, int target_layout);          // -> kernel layout id (0 = 12/12), -1 if unsupported
int svb_check_layout_form(int lay, bool mixed, bool has_grad);
int svb_launch_finalize(const float* part_render, const float* part_l1, int B, int HW, bool packed, int N, bool mixed,
                        float l1_weight, float* out, int n_out, cudaStream_t st);
