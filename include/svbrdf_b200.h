/*
 * svbrdf_b200.h - C ABI of the B200-native rendering-loss path.
 *
 * Drop-in boundary for the one hot path of mworchel/svbrdf-estimation that this library
 * replaces (paths relative to development/multiImage_pytorch/ of the reference):
 *
 *   LocalRenderer.render(scene, svbrdf)         renderers.py:67-104
 *   RenderingLoss.forward(input, target)        losses.py:29-52
 *   MixedLoss / SVBRDFL1Loss                    losses.py:7-19, 54-63   (caller of the path)
 *   Camera / Light / Scene                      environment.py:4-16     (the "scene record")
 *
 * The reference has no FFI of its own (it is eager PyTorch); the host side that binds
 * these symbols is svbrdf_estimation_b200/_cabi.py (ctypes), and INTEGRATION.md shows the
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain pointers and sizes only.  All "dev" pointers are device memory of the CURRENT
 *     CUDA device, fp32, contiguous NCHW, 4-byte aligned (8-byte aligned pointers and an even W
 *     select the two-pixels-per-thread kernels; anything else runs the one-pixel-per-thread ones);
 *     "host" pointers are ordinary host memory.
 *   - The caller owns every buffer.  The library allocates nothing in the device-pointer
 *     entry points, keeps no state between calls, and is re-entrant.  (The *_host entry
 *     points use an explicit context object that owns staging buffers and streams.)
 *   - All work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream).  Nothing synchronises unless documented.
 *   - Return value: 0 on success, otherwise a negative SVBRDF_E_* code or a positive
 *     cudaError_t; svbrdf_b200_last_error() gives a thread-local message.  Nothing throws.
 *   - Maps are [B,12,H,W] with channels 0-2 normals, 3-5 diffuse, 6-8 roughness (one per
 *     colour channel), 9-11 specular (utils.py:36-58).  H must equal W (renderers.py:73-76).
 *   - A scene record is 9 floats: camera xyz, light xyz, light colour rgb.  Scene records
 *     are HOST memory: they travel to the device as kernel parameters (constant bank),
 *     so there is no scene upload and no device-side scene buffer.
 *   - `lin` is the W-entry coordinate table torch.linspace(-1, 1, W) (renderers.py:73);
 *     pixel (row, col) sits at (lin[col], -lin[row], 0).
 */
#ifndef SVBRDF_B200_H
#define SVBRDF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVBRDF_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define SVBRDF_API __attribute__((visibility("default")))
#else
#define SVBRDF_API
#endif

#define SVBRDF_E_INVALID   (-1) /* bad argument (null pointer, non-positive size, H != W ...) */
#define SVBRDF_E_TOO_LARGE (-2) /* size exceeds what the index arithmetic supports            */
#define SVBRDF_E_STATE     (-3) /* context misuse in the *_host entry points                  */

SVBRDF_API int svbrdf_b200_abi_version(void);

/* Thread-local description of the last non-zero status returned on this thread. */
SVBRDF_API const char* svbrdf_b200_last_error(void);

/* Hash of the CUDA sources this binary was built from (svbrdf_estimation_b200/_build.py); the Python binding
 * compares it with the sources on disk and refuses (or rebuilds) a stale library.                       */
SVBRDF_API const char* svbrdf_b200_build_id(void);

/* Bytes of device workspace the loss entry points need for a [B,12,H,W] problem with N
 * scene records per batch element (per-CTA loss partials).  Never 0.                      */
SVBRDF_API size_t svbrdf_b200_workspace_bytes(int B, int N, int H, int W);

/* Fills lin_host[0..W) with torch.linspace(-1, 1, W) (renderers.py:73), bit-identical to torch's CPU and
 * CUDA results; upload it once per W and pass it as `lin_dev` below.  Host-only, no CUDA call.       */
SVBRDF_API int svbrdf_b200_coordinate_table(float* lin_host, int W);

/* ---- scene sampling (environment.py:18-55), host side ---------------------------------------------
 * Writes records_host[B][n_random + n_specular][9] for batch elements first_batch_element ..
 * first_batch_element + B - 1: n_random configurations with independently cosine-sampled view and light
 * directions at unit distance (colour 20) followed by n_specular mirror configurations with
 * log-normal distances and a common xy shift (colour 50).  Stateless and counter-based: the scenes of
 * batch element e depend only on (seed, e), so ranks that own different batch slices draw disjoint,
 * reproducible scenes.  Same distributions as the reference, different random stream (the Python
 * samplers in environment.py reproduce the reference's torch draws).  Host-only, no CUDA call.       */
SVBRDF_API int svbrdf_b200_sample_scenes(uint64_t seed, int first_batch_element, int B, int n_random,
                                         int n_specular, float* records_host);

/* The random draws of RenderingLoss.forward's samplers for B batch elements (losses.py:35: generate_random_scenes(
 * n_random) + generate_specular_scenes(n_specular) per element; environment.py:18-55, utils.py:100-111), taken from
 * torch's global CPU generator in the reference's order.  torch_cpu_rng_state is the buffer of torch.get_rng_state()
 * (5056 bytes: at::CPUGeneratorImplState); it is advanced in place - hand it to torch.set_rng_state() afterwards and
 * torch continues exactly where the reference's own calls would have left it.
 *   uniforms_host [B][4*n_random + 4*n_specular]: raw uniform_(0,1) draws per element in stream order
 *                 (view r1.., r2.., light r1.., r2.. | specular view r1.., r2.. | shift x0,y0,x1,y1..)
 *   normals_host  [B][2][n_specular]: the normal_(0.5, 0.75) log-distances (view, light)
 * n_specular must be < 16 (normal_() on larger tensors takes ATen's vectorised path).  Host-only, no CUDA call.  */
SVBRDF_API int svbrdf_b200_reference_draws(void* torch_cpu_rng_state, size_t state_bytes, int B, int n_random,
                                int n_specular, float* uniforms_host, float* normals_host);

/* The same sampling in the two-phase form RenderingLoss uses: everything of environment.py:18-55 / utils.py:100-111
 * except sqrt, cos, sin and exp (torch's vectorised / MKL implementations, applied by the caller to whole-batch
 * tensors between the two calls) - draws, uniform_(lo, hi) maps, products and sums, each rounded to float where the
 * reference's float32 tensor ops round.  Directions per batch element: n_random views, n_random lights, n_specular
 * mirror views.
 *   begin : r1_host, phi_host [B][2*n_random + n_specular]; log_distance_host [B][2][n_specular];
 *           shift_host [B][n_specular][2]; advances torch_cpu_rng_state like svbrdf_b200_reference_draws
 *   finish: radius_host = torch.sqrt(r1), cos_phi_host / sin_phi_host = torch.cos / torch.sin(phi),
 *           z_host = torch.sqrt(1 - radius * radius), distance_host = torch.exp(log_distance)
 *           ->  records_host [B][n_random + n_specular][9], bit-identical to the reference's scenes            */
SVBRDF_API int svbrdf_b200_reference_scenes_begin(void* torch_cpu_rng_state, size_t state_bytes, int B, int n_random,
                                int n_specular, float* r1_host, float* phi_host,
                                float* log_distance_host, float* shift_host);
SVBRDF_API int svbrdf_b200_reference_scenes_finish(int B, int n_random, int n_specular, const float* radius_host,
                                const float* cos_phi_host, const float* sin_phi_host, const float* z_host,
                                const float* distance_host, const float* shift_host, float* records_host);

/* ---- LocalRenderer.render (renderers.py:67-104) ------------------------------------------
 * images[b,k,:,:,:] = radiance of maps[b] under scene record k of batch element b.
 *   scenes_host : [B,N,9] if scenes_per_batch != 0, else [N,9] shared by every b
 *                 (render(scene, svbrdf) applies ONE scene to the whole batch).
 *   images_dev  : [B,N,3,H,W], linear radiance, unclamped.                                   */
SVBRDF_API int svbrdf_b200_render_forward(const float* maps_dev, int B, int H, int W,
                               const float* scenes_host, int N, int scenes_per_batch,
                               const float* lin_dev, float* images_dev, void* stream);

/* Vector-Jacobian product of the above w.r.t. the maps (autograd of renderers.py:67-104):
 * grad_maps[b,:,:,:] = sum_k J^T grad_images[b,k].  grad_maps_dev is overwritten.          */
SVBRDF_API int svbrdf_b200_render_backward(const float* maps_dev, int B, int H, int W,
                                const float* scenes_host, int N, int scenes_per_batch,
                                const float* lin_dev, const float* grad_images_dev,
                                float* grad_maps_dev, void* stream);

/* ---- RenderingLoss.forward (losses.py:29-52) ------------------------------------------------
 * loss = mean_{b,k,c,y,x} | log(R(input)+0.1) - log(R(target)+0.1) |, written as one float
 * to loss_dev.  scenes_host is [B,N,9] (fresh scenes per batch element, losses.py:35).
 * Deterministic: per-CTA partials are summed in a fixed order in fp64.                       */
SVBRDF_API int svbrdf_b200_loss_forward(const float* input_dev, const float* target_dev, int B, int H, int W,
                             const float* scenes_host, int N, const float* lin_dev,
                             float* loss_dev, void* workspace_dev, size_t workspace_bytes,
                             void* stream);

/* Same pass that also writes grad_input_dev[B,12,H,W] = d loss / d input (the target never
 * requires grad in the reference's callers, main.py:111,116).  The gradient is for an
 * upstream gradient of 1; see svbrdf_b200_scale_grad.                                        */
SVBRDF_API int svbrdf_b200_loss_forward_backward(const float* input_dev, const float* target_dev,
                                      int B, int H, int W, const float* scenes_host, int N,
                                      const float* lin_dev, float* loss_dev, float* grad_input_dev,
                                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same call with the accurate-highlight evaluation of the GGX denominator that render_forward always uses
 * (1 - (n.h)^2 from |n x (wi+wo)|^2): loss and gradient 20-30x closer to an fp64 evaluation
 * of renderers.py:67-104 than the reference's own fp32 run, about 12 % lower throughput.                     */
SVBRDF_API int svbrdf_b200_loss_forward_backward_accurate(const float* input_dev, const float* target_dev,
                                      int B, int H, int W, const float* scenes_host, int N,
                                      const float* lin_dev, float* loss_dev, float* grad_input_dev,
                                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* Forward-only form of the accurate-highlight evaluation (no gradient buffer).                              */
SVBRDF_API int svbrdf_b200_loss_forward_accurate(const float* input_dev, const float* target_dev, int B, int H, int W,
                             const float* scenes_host, int N, const float* lin_dev,
                             float* loss_dev, void* workspace_dev, size_t workspace_bytes,
                             void* stream);

/* grad[i] *= *upstream_dev for i < count, in place; returns immediately on the device when
 * *upstream_dev == 1.0f (the loss.backward() case), so no host synchronisation is needed
 * to skip the pass.                                                                          */
SVBRDF_API int svbrdf_b200_scale_grad(float* grad_dev, size_t count, const float* upstream_dev, void* stream);

/* ---- MixedLoss (losses.py:54-63): l1_weight * SVBRDFL1Loss + RenderingLoss in one pass ------
 * out_dev[0] = mixed loss, out_dev[1] = rendering loss, out_dev[2] = map L1 loss (unweighted).
 * grad_input_dev may be NULL (forward only).                                                 */
SVBRDF_API int svbrdf_b200_mixed_loss_forward_backward(const float* input_dev, const float* target_dev,
                                            int B, int H, int W, const float* scenes_host, int N,
                                            float l1_weight, const float* lin_dev, float* out_dev,
                                            float* grad_input_dev, void* workspace_dev,
                                            size_t workspace_bytes, void* stream);

/* Same, fed by the network's ENCODED output (models.py:334-346 + utils.py:73-98 fused in): encoded_dev is
 * [B,9,H,W] in [-1,1] after tanh - normal xy, diffuse rgb, roughness, specular rgb.  The kernel decodes
 * it to the 12-channel maps on the fly (n = normalize(3x,3y,1); (v+1)/2 for diffuse/roughness/specular;
 * roughness replicated) and writes grad_encoded_dev[B,9,H,W] = d mixed loss / d encoded.             */
SVBRDF_API int svbrdf_b200_mixed_loss_encoded_forward_backward(const float* encoded_dev, const float* target_dev,
                                            int B, int H, int W, const float* scenes_host, int N,
                                            float l1_weight, const float* lin_dev, float* out_dev,
                                            float* grad_encoded_dev, void* workspace_dev,
                                            size_t workspace_bytes, void* stream);

/* ---- channel layouts -------------------------------------------------------------------------------------
 * The reference's tensors carry 12 channels of which the three roughness channels are replicas of one map
 * (utils.py:78-80: the model and the dataset repeat it), and the network itself emits 9 (models.py:334-346).  Callers
 * that know this can hand over fewer bytes; the gradient comes back in the input's layout.
 *   SVBRDF_LAYOUT_MAPS12   [B,12,H,W]  normals(3) diffuse(3) roughness(3) specular(3)          utils.py:36-58
 *   SVBRDF_LAYOUT_MAPS10   [B,10,H,W]  normals(3) diffuse(3) roughness(1) specular(3); d/d roughness = sum of the three
 *   SVBRDF_LAYOUT_ENCODED9 [B, 9,H,W]  network output after tanh: normal xy, diffuse, roughness, specular in [-1,1]
 * Supported (input, target): (MAPS12, MAPS12), (MAPS10, MAPS10), (ENCODED9, MAPS12), (ENCODED9, MAPS10).
 * l1_weight < 0: RenderingLoss (losses.py:29-52); l1_weight >= 0: MixedLoss (losses.py:54-63).  ENCODED9 input exists as
 * MixedLoss forward+backward only, MAPS10 input as RenderingLoss only.  out_dev[0..2] = loss, rendering loss, map-L1 loss;
 * grad_input_dev (input layout) may be NULL where a forward-only form exists.                                   */
#define SVBRDF_LAYOUT_MAPS12   12
#define SVBRDF_LAYOUT_MAPS10   10
#define SVBRDF_LAYOUT_ENCODED9 9
SVBRDF_API int svbrdf_b200_loss_layouts(const float* input_dev, int input_layout, const float* target_dev, int target_layout,
                                        int B, int H, int W, const float* scenes_host, int N, float l1_weight,
                                        const float* lin_dev, float* out_dev, float* grad_input_dev, void* workspace_dev,
                                        size_t workspace_bytes, void* stream);

/* ---- host-buffer entry point (the call a non-PyTorch caller makes) --------------------------
 * A context owns pinned staging buffers, device buffers and copy/compute streams for problems
 * up to the given size on the current device.  One call at a time per context (use one context per
 * host thread); different contexts are independent.                                          */
typedef struct svbrdf_b200_ctx svbrdf_b200_ctx;

SVBRDF_API int svbrdf_b200_ctx_create(svbrdf_b200_ctx** out, int max_B, int max_N, int H, int W);
SVBRDF_API void svbrdf_b200_ctx_destroy(svbrdf_b200_ctx* ctx);

/* Pinned host buffers owned by the context (input, target, grad: max_B*12*H*W floats each).
 * Filling these directly avoids an extra host-side copy.  which: 0 input, 1 target, 2 grad.  */
SVBRDF_API float* svbrdf_b200_ctx_pinned(svbrdf_b200_ctx* ctx, int which);

/* RenderingLoss forward+backward on HOST maps: uploads input/target (batch-chunked, copies
 * overlapped with the kernels), computes, downloads grad_input and the loss.  input_host /
 * target_host / grad_host may be the context's pinned buffers or any host memory (pageable
 * memory is staged through the pinned buffers).  Blocks until the results are on the host.  */
SVBRDF_API int svbrdf_b200_rendering_loss_host(svbrdf_b200_ctx* ctx, const float* input_host,
                                    const float* target_host, int B, const float* scenes_host,
                                    int N, float* loss_host, float* grad_host);

/* The same pipeline for any supported layout pair and for MixedLoss (see svbrdf_b200_loss_layouts): with MAPS10 maps a
 * step moves 30 instead of 36 planes over PCIe, with ENCODED9 input and a MAPS10 target 28.  out_host[0..2] = loss,
 * rendering loss, map-L1 loss; grad_host has the input's layout (NULL: no gradient, where that form exists).       */
SVBRDF_API int svbrdf_b200_loss_host(svbrdf_b200_ctx* ctx, const float* input_host, int input_layout,
                                    const float* target_host, int target_layout, int B, const float* scenes_host,
                                    int N, float l1_weight, float* out_host, float* grad_host);

#ifdef __cplusplus
}
#endif
#endif /* SVBRDF_B200_H */
