/*
 * vxm_b200 — C ABI of the B200-native VxmDense registration path.
 *
 * Drop-in boundary.  The reference (voxelmorph/voxelmorph, torch backend) has no FFI
 * layer of its own: its hot path bottoms out in PyTorch operators (F.grid_sample,
 * F.interpolate, nn.Conv3d, F.conv3d, torch.optim.Adam).  Each entry point below
 * replaces one of those call sites; the reference file:line it stands in for is cited
 * on the declaration.  The host side that binds them (ctypes) is
 * voxelmorph_b200/_lib.py; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only: every pointer is a DEVICE pointer owned by the caller
 *    (PyTorch caching allocator); the library allocates nothing persistent;
 *  - tensors are contiguous, float32, batch-major "NCDHW" (B, C, D, H, W) unless a
 *    declaration says otherwise; a 2-D problem is passed with D == 1 and nd == 2
 *    (flows then carry 2 channels: H, W);
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit
 *    synchronisation; the current CUDA device is the caller's;
 *  - return value 0 on success, negative on error; vxm_last_error() gives the text
 *    (thread local).  Nothing throws or aborts;
 *  - `*_workspace_bytes` functions are pure host arithmetic (no CUDA calls).
 */
#ifndef VXM_B200_H
#define VXM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXM_OK 0
#define VXM_ERR_ARG (-1)
#define VXM_ERR_CUDA (-2)
#define VXM_ERR_UNSUPPORTED (-3)

/* interpolation mode of the resampler (reference SpatialTransformer(mode=...), layers.py:11-14) */
#define VXM_MODE_LINEAR 0
#define VXM_MODE_NEAREST 1
/* how `loc / (S-1)` of layers.py:37 is rounded: a true fp32 division (torch CPU) or a
 * multiply by fl(1/(S-1)) (what torch's CUDA `tensor / python_scalar` computes). */
#define VXM_ARITH_TRUE_DIV 0
#define VXM_ARITH_RECIPROCAL 1
/* linear mode only: coord = (p + flow) * (Ssrc-1)/(S-1) without replaying the reference's fp32 round trip; results
 * agree with the exact modes to a few 1e-6 of the value range (north_star tolerance for floating point: 1e-4) and the
 * kernels are memory bound instead of instruction bound.  The nearest mode always replays the exact arithmetic. */
#define VXM_ARITH_FAST 2

const char* vxm_last_error(void);
/* library / build identification ("vxm_b200 <version> sm_100a") */
const char* vxm_version(void);
/* number of kernel launches issued through this library by the calling process so far */
uint64_t vxm_launch_count(void);

/* ---- SpatialTransformer: reference voxelmorph/torch/layers.py:30-48 (F.grid_sample :48) ----
 * out[b,c,p] = sample(src[b,c], p + flow[b,:,p]); zeros padding; align_corners=True.
 * src: (B,C,Ds,Hs,Ws)  flow: (B,nd,D,H,W)  out: (B,C,D,H,W).  The reference always uses
 * (Ds,Hs,Ws) == (D,H,W); distinct sizes follow grid_sample's semantics. */
int vxm_warp_fwd(const float* src, const float* flow, float* out,
                 int B, int C, int Ds, int Hs, int Ws, int D, int H, int W, int nd,
                 int mode, int arith, void* stream);
/* backward of the above.  grad_src (B,C,Ds,Hs,Ws) is ACCUMULATED into (caller zero-fills) and
 * may be NULL; grad_flow (B,nd,D,H,W) is overwritten and may be NULL. */
int vxm_warp_bwd(const float* grad_out, const float* src, const float* flow,
                 float* grad_src, float* grad_flow,
                 int B, int C, int Ds, int Hs, int Ws, int D, int H, int W, int nd,
                 int mode, int arith, void* stream);

/* ---- VecInt: reference voxelmorph/torch/layers.py:51-68 (scaling and squaring) ----
 * vel, out: (B,nd,D,H,W).  All nsteps squarings run in one cooperative launch.
 * If `states` is non-NULL it receives the nsteps intermediate fields v_0..v_{n-1}
 * (nsteps * B*nd*D*H*W floats) needed by the backward; otherwise `work` must hold
 * vxm_vecint_workspace_bytes() bytes of scratch. */
size_t vxm_vecint_workspace_bytes(int B, int D, int H, int W, int nd, int nsteps);
int vxm_vecint_fwd(const float* vel, float* out, float* states, void* work,
                   int B, int D, int H, int W, int nd, int nsteps, int arith, void* stream);
/* grad_vel (B,nd,D,H,W) overwritten.  `states` as written by the forward; `work` holds
 * 2 * B*nd*D*H*W floats of scratch. */
int vxm_vecint_bwd(const float* grad_out, const float* states, float* grad_vel, void* work,
                   int B, int D, int H, int W, int nd, int nsteps, int arith, void* stream);
/* arith == VXM_ARITH_FAST (3-D, nsteps >= 1): the launch keeps the field in an interleaved float4 (z,y,x,0) layout.
 * `states` then holds vxm_vecint_fast_states_bytes() bytes (nsteps float4 fields), `work`
 * vxm_vecint_fast_work_bytes(backward) bytes (2 float4 fields forward without states, 3 backward). */
size_t vxm_vecint_fast_states_bytes(int B, int D, int H, int W, int nsteps);
size_t vxm_vecint_fast_work_bytes(int B, int D, int H, int W, int backward);
/* measurement aid: a cooperative launch (512 threads per CTA, `ctas_per_sm` CTAs per SM or the occupancy limit when 0)
 * that executes `nsync` grid-wide synchronisations and nothing else */
int vxm_debug_gridsync(int nsync, int ctas_per_sm, void* stream);

/* ---- ResizeTransform: reference voxelmorph/torch/layers.py:85-97 (F.interpolate :88,:94) ----
 * out = post * lerp(pre * x) with align_corners=True linear interpolation.
 * x: (B,C,Di,Hi,Wi)  out: (B,C,Do,Ho,Wo). */
int vxm_resize_fwd(const float* x, float* out, int B, int C, int Di, int Hi, int Wi,
                   int Do, int Ho, int Wo, float pre, float post, void* stream);
/* adjoint (deterministic gather form).  grad_x overwritten. */
int vxm_resize_bwd(const float* grad_out, float* grad_x, int B, int C, int Di, int Hi, int Wi,
                   int Do, int Ho, int Wo, float pre, float post, void* stream);

/* ---- NCC: reference voxelmorph/torch/losses.py:15-67 (5 x F.conv3d with a ones filter) ----
 * I = y_true, J = y_pred: (B,1,D,H,W).  win = (wd,wh,ww) odd window (1 along D for 2-D).
 * loss[0] = -mean(cc).  `work`: vxm_ncc_workspace_bytes().  If `saved` is non-NULL the
 * forward stores 4 fields (4 * B*D*H*W floats) that the backward consumes. */
size_t vxm_ncc_workspace_bytes(int B, int D, int H, int W);
int vxm_ncc_fwd(const float* I, const float* J, float* loss, float* saved, void* work,
                int B, int D, int H, int W, int wd, int wh, int ww, void* stream);
/* grad_J = grad_loss[0] * d(-mean cc)/dJ.  grad_loss is a device scalar. */
int vxm_ncc_bwd(const float* I, const float* J, const float* saved, const float* grad_loss,
                float* grad_J, int B, int D, int H, int W, int wd, int wh, int ww, void* stream);

/* ---- Jacobian determinant of x -> x + disp(x): reference voxelmorph/py/utils.py:473-516 (numpy, np.gradient) ----
 * disp: (B,nd,D,H,W) displacement in voxels (the layout VxmDense(registration=True) returns).  det (B,D,H,W), may be NULL;
 * folds (one uint64, may be NULL) receives the number of voxels with det <= 0. */
int vxm_jacdet(const float* disp, float* det, unsigned long long* folds, int B, int D, int H, int W, int nd, void* stream);

/* ---- Grad: reference voxelmorph/torch/losses.py:102-135 ----
 * y: (B,nd,D,H,W) (any channel count C).  penalty 1 = l1, 2 = l2.  loss[0] = mult * mean_b mean_axes mean |dy|^p */
size_t vxm_reduce_workspace_bytes(void);
int vxm_gradloss_fwd(const float* y, float* loss, void* work, int B, int C, int D, int H, int W,
                     int nd, int penalty, float mult, void* stream);
int vxm_gradloss_bwd(const float* y, const float* grad_loss, float* grad_y, int B, int C, int D,
                     int H, int W, int nd, int penalty, float mult, void* stream);

/* ---- MSE: reference voxelmorph/torch/losses.py:75-76 ---- */
int vxm_mse_fwd(const float* y_true, const float* y_pred, float* loss, void* work, size_t n,
                void* stream);
int vxm_mse_bwd(const float* y_true, const float* y_pred, const float* grad_loss,
                float* grad_pred, size_t n, void* stream);

/* ---- Dice: reference voxelmorph/torch/losses.py:84-90 ----
 * y_true, y_pred: (B,L,V) with V = D*H*W.  `work`: vxm_dice_workspace_bytes(B*L).
 * `sums` (2*B*L floats: top, bottom(unclamped)) is written for the backward. */
size_t vxm_dice_workspace_bytes(int BL);
int vxm_dice_fwd(const float* y_true, const float* y_pred, float* loss, float* sums, void* work,
                 int BL, size_t V, void* stream);
int vxm_dice_bwd(const float* y_true, const float* sums, const float* grad_loss, float* grad_pred,
                 int BL, size_t V, void* stream);

/* ---- Conv3d k=3 s=1 p=1 (+bias, +LeakyReLU 0.2): reference voxelmorph/torch/networks.py:299-304
 * (ConvBlock), :211,:257 (flow head, no activation).  fp32 "parity" engine, NCDHW.
 * x: (B,Cin,D,H,W)  w: (Cout,Cin,kd,3,3) with kd = 3 (3-D) or 1 (2-D)  y: (B,Cout,D,H,W).
 * leaky_slope < 0 disables the activation. */
int vxm_conv3d_fwd_f32(const float* x, const float* w, const float* bias, float* y,
                       int B, int Cin, int Cout, int D, int H, int W, int kd,
                       float leaky_slope, void* stream);
/* grad_y is the gradient w.r.t. the ACTIVATED output y; y (saved forward output) supplies the
 * LeakyReLU mask (y < 0).  grad_x may be NULL (first layer).  grad_w / grad_b are ACCUMULATED
 * into (caller zero-fills once per step).  work: vxm_conv3d_bwd_workspace_bytes(). */
size_t vxm_conv3d_bwd_workspace_bytes(int B, int Cin, int Cout, int D, int H, int W, int kd);
int vxm_conv3d_bwd_f32(const float* grad_y, const float* y, const float* x, const float* w,
                       float* grad_x, float* grad_w, float* grad_b, void* work,
                       int B, int Cin, int Cout, int D, int H, int W, int kd,
                       float leaky_slope, void* stream);

/* ---- Conv3d k=3 on tcgen05 tensor cores (bf16 operands, fp32 TMEM accumulation), channels-last ----
 * The throughput engine for reference networks.py:299-304 / :211,257; forward and dgrad share one kernel.
 * Activations are bf16 NDHWC (B,D,H,W,C).  Weights are pre-packed by vxm_conv3d_tc_pack into the UMMA
 * canonical K-major layout [tap][K/16][2][N][8] (bf16); `transposed` = 1 packs the dgrad operator
 * (channel roles swapped, taps flipped).  np = MMA N (16 or 32) >= number of output channels. */
size_t vxm_conv3d_tc_packed_bytes(int cin_eff, int np, int kd);
int vxm_conv3d_tc_pack(const float* w, void* wpk, int Cout, int Cin, int kd, int np, int transposed,
                       void* stream);
/* Input = channel concat of [xa (Ca ch; at HALF resolution and nearest-upsampled x2 on the fly when up=1),
 * xb (Cb ch)], both bf16 NDHWC; or, when nplanar > 0, `nplanar` (<= 4) planar fp32 volumes xf[i]
 * ((B,D,H,W) each, batch stride xf_bstride[i] floats) converted on load (first layer: source/target images;
 * flow-head dgrad: the 3 flow-gradient planes).  out_mode 0: bf16 NDHWC (B,D,H,W,Cout), Cout % 8 == 0;
 * out_mode 1: fp32 NCDHW.  Epilogue: + bias (may be NULL), then LeakyReLU(slope) if slope >= 0; if `mask`
 * (bf16 NDHWC, same shape as out) is given the activation is replaced by  out *= (mask < 0 ? slope : 1)
 * (LeakyReLU derivative of the layer below, for dgrad). */
int vxm_conv3d_tc_fwd(const void* xa, const void* xb, const float* const* xf, const long long* xf_bstride,
                      int nplanar, const void* wpk, const float* bias, void* out, const void* mask,
                      int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int np, int kd,
                      int out_mode, float slope, void* out2, int csplit, void* stream);
/* out2 != NULL (bf16 outputs only): channels [0,csplit) go to `out` (B,D,H,W,csplit) and [csplit,Cout) to `out2`
 * (B,D,H,W,Cout-csplit) — the single-pass dgrad of a layer whose input was a channel concat; np may then be 48 or 64. */
/* "kw-stacked" variant of the tensor-core convolution (Cin in {8,16,32,48,64}, Cout <= 64): the three kw taps are
 * stacked along the MMA N dimension (N = 3*coutp), the K loop runs over (kd,kh,Cin/16) only and the kw shift is a warp
 * shuffle in the epilogue — the activation operand is read 9x instead of 27x.  Same inputs / outputs / epilogue options
 * as vxm_conv3d_tc_fwd (no planar sources); weights are packed by vxm_conv3d_tct_pack (coutp in {16,32,48,64}). */
size_t vxm_conv3d_tct_packed_bytes(int cin_eff, int coutp, int kd);
int vxm_conv3d_tct_pack(const float* w, void* wpk, int Cout, int Cin, int kd, int coutp, int transposed, void* stream);
int vxm_conv3d_tct_supported(int Ca, int Cb, int Cout);
int vxm_conv3d_tct_fwd(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                       int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                       float slope, void* out2, int csplit, void* stream);
/* Swizzled-operand variant of the kw-stacked kernel (128/64/32-byte swizzled K-major shared-memory layouts): same
 * arguments and semantics as vxm_conv3d_tct_*; weights are packed by vxm_conv3d_tcs_pack. */
size_t vxm_conv3d_tcs_packed_bytes(int cin_eff, int coutp, int kd);
int vxm_conv3d_tcs_pack(const float* w, void* wpk, int Cout, int Cin, int kd, int coutp, int transposed, void* stream);
int vxm_conv3d_tcs_supported(int Ca, int Cb, int Cout);
/* every packed operand of a model in ONE launch: the caller fills an array of descriptors on the host
 * (vxm_conv3d_tcs_pack_desc_bytes() bytes each; vxm_conv3d_tcs_pack_desc returns the operand's element count so that
 * `begin` can be chained), uploads it once and calls vxm_conv3d_tcs_pack_multi with the summed element count. */
size_t vxm_conv3d_tcs_pack_desc_bytes(void);
int vxm_conv3d_tcs_pack_desc(void* desc_host, const float* w, void* wpk, int Cout, int Cin, int kd, int coutp, int transposed,
                             int begin);
/* descriptor of a kd-folded 2-D operand of the 3-D weight w (Cout, Cin, 3, 3, 3): operand input channel kd * r + c is tap kd of
 * real channel c, r = Cin (Cout when transposed), 3 r <= 16; packed size = vxm_conv3d_tcs_packed_bytes(3 r, coutp, 1) */
int vxm_conv3d_tcs_pack_desc_fold(void* desc_host, const float* w, void* wpk, int Cout, int Cin, int coutp, int transposed, int begin);
int vxm_conv3d_tcs_pack_multi(const void* descs_dev, int ndesc, int total, void* stream);
int vxm_conv3d_tcs_fwd(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                       int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                       float slope, void* out2, int csplit, void* stream);
/* variant of the above with two alternating MMA-issuing warps (8-row tiles only: Ca + Cb in {8,16,32,48}, padded
 * Cout in {16,32}); same arguments, weights packed by vxm_conv3d_tcs_pack.  Selected with VXM_B200_TCS2=1. */
int vxm_conv3d_tcs2_fwd(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, const void* mask,
                       int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp, int kd, int out_mode,
                       float slope, void* out2, int csplit, void* stream);
/* Split-precision ("bf16x3") passes of the same kernel — the in-tolerance tensor-core mode (reference layer:
 * voxelmorph/torch/networks.py:290-305 in fp32).  Every operand is a bf16 pair hi + lo (16 mantissa bits); a layer is
 * three launches that accumulate  x_lo*w_hi + x_hi*w_lo + x_hi*w_hi  in fp32:
 *   out_mode 2: out (fp32, channels-last, `coutp` channels per voxel) = acc_in + conv   (no bias / activation; acc_in may
 *               be NULL or alias out);
 *   out_mode 3: x = act(acc_in + conv + bias) stored as the bf16 pair out (hi), out_lo (lo = bf16(x - hi));
 *   out_mode 1: fp32 planar out = acc_in + conv + bias (the flow head).
 * acc_in has `coutp` channels per voxel. */
int vxm_conv3d_tcs_fwd_acc(const void* xa, const void* xb, const void* wpk, const float* bias, void* out, void* out_lo,
                           const float* acc_in, int B, int D, int H, int W, int Ca, int Cb, int up, int Cout, int coutp,
                           int kd, int out_mode, float slope, void* stream);
/* Weight (and bias) gradient on tensor cores.  x sources as in vxm_conv3d_tc_fwd (the layer's forward input);
 * gz = gradient w.r.t. the convolution output (already multiplied by the activation derivative): bf16 NDHWC with
 * Cg in {8,16,32} channels, or nplanar_g (<= 4) planar fp32 volumes (flow head).  grad_w: fp32
 * (Cout_real, Cin_real, kd, 3, 3); grad_b: fp32 (Cout_real), may be NULL.  accumulate = 0 overwrites them, 1 adds to
 * them (gradient buffers zeroed once per step).  work: vxm_conv3d_tc_wgrad_workspace_bytes(kd). */
size_t vxm_conv3d_tc_wgrad_workspace_bytes(int kd);
int vxm_conv3d_tc_wgrad(const void* xa, const void* xb, const float* const* xf, const long long* xf_bstride,
                        int nplanar_x, const void* gz, const float* const* gf, const long long* gf_bstride,
                        int nplanar_g, float* grad_w, float* grad_b, void* work, int B, int D, int H, int W,
                        int Ca, int Cb, int up, int Cin_real, int Cg, int Cout_real, int kd, int accumulate, void* stream);
/* Deferred reduction of the weight gradient (channels-last bf16 sources only): `_partial` launches the tcgen05 kernel(s) of
 * one layer into caller-provided workspace (`work_used` bytes of it are then owned by this layer until the flush) and
 * appends the pending reductions to a HOST array of descriptors (vxm_conv3d_tc_wgrad2_desc_bytes() each, at most
 * vxm_conv3d_tc_wgrad2_max_pending()); `_flush` reduces every pending layer in ONE launch, in a fixed order
 * (deterministic).  Arguments as vxm_conv3d_tc_wgrad. */
size_t vxm_conv3d_tc_wgrad2_desc_bytes(void);
int vxm_conv3d_tc_wgrad2_max_pending(void);
size_t vxm_conv3d_tc_wgrad2_partial_bytes(int kd);
int vxm_conv3d_tc_wgrad2_partial(const void* xa, const void* xb, const void* gz, float* grad_w, float* grad_b, void* work,
                                 size_t work_bytes, size_t* work_used, void* descs_host, int* ndesc, int B, int D, int H, int W,
                                 int Ca, int Cb, int up, int Cin_real, int Cg, int Cout_real, int kd, int accumulate, void* stream);
int vxm_conv3d_tc_wgrad2_flush(const void* descs_host, int ndesc, void* stream);
/* Weight gradient of a "kd-folded" layer (vxm_planar_fold_kd_bf16): x (B,D,H,W,Cx) against gz (B,D,H,W,Cg), Cx, Cg in {8,16}, as
 * a 2-D problem per slice with the kh taps stacked in the MMA's M; grad_w has the 2-D layout (Cout_real, Cin_real, 1, 3, 3).
 * Deferred like vxm_conv3d_tc_wgrad2_partial (one pending reduction). */
int vxm_conv3d_tc_wgrad2_partial_khm(const void* x, const void* gz, float* grad_w, float* grad_b, void* work, size_t work_bytes,
                                     size_t* work_used, void* descs_host, int* ndesc, int B, int D, int H, int W, int Cx,
                                     int Cin_real, int Cg, int Cout_real, int accumulate, void* stream);
/* ---- channels-last bf16 glue of the tensor-core U-Net engine (reference networks.py:126-138 and its autograd) ----
 * All tensors bf16 (B,D,H,W,C), C % 8 == 0.  (Dc,Hc,Wc) are the COARSE dims; the fine tensor is (fd*Dc, 2Hc, 2Wc)
 * with fd = 2 for nd == 3 and 1 for nd == 2. */
int vxm_pool2_ndhwc_bf16(const void* x_fine, void* y_coarse, int B, int Dc, int Hc, int Wc, int C, int nd, void* stream);
/* out_coarse = (sum over the 2^nd children of g_fine) * (act_coarse < 0 ? slope : 1); act_coarse may be NULL */
int vxm_sumpool_mask_ndhwc_bf16(const void* g_fine, const void* act_coarse, void* out_coarse, int B, int Dc, int Hc,
                                int Wc, int C, int nd, float slope, void* stream);
/* out_fine = (g_skip_fine + [child is the first max of e_fine in its window] * g_pool_coarse) * (e_fine < 0 ? slope : 1);
 * g_skip or g_pool may be NULL (not both) */
int vxm_unpool_combine_ndhwc_bf16(const void* e_fine, const void* g_skip_fine, const void* g_pool_coarse, void* out_fine,
                                  int B, int Dc, int Hc, int Wc, int C, int nd, float slope, void* stream);
/* out (B,V,8) bf16 <- up to 8 planar fp32 volumes (channel c = planes[c], batch stride bstrides[c] floats); unused
 * channels are zero.  Feeds the fp32 images / the fp32 flow gradient to the tensor-core kernels. */
int vxm_planar_to_ndhwc8_bf16(const float* const* planes, const long long* bstrides, int nplanes, void* out, int B,
                              size_t V, void* stream);
/* kd folded into the channels: out (B,D,HW,cout) bf16, cout in {8,16}, channel kd * nplanes + p = planes[p] at slice d + kd - 1
 * (zero outside the volume), kd = 0..2; 3 * nplanes <= cout.  A 3-D convolution with so few real input channels then runs as a
 * 2-D one over the folded tensor (reference layers: the first ConvBlock, networks.py:122-130, and the flow head's autograd). */
int vxm_planar_fold_kd_bf16(const float* const* planes, const long long* bstrides, int nplanes, void* out, int B, int D,
                            size_t HW, int cout, void* stream);
/* split-precision variants: out_hi = bf16(x), out_lo = bf16(x - out_hi); MaxPool(2) of a (hi, lo) pair tensor (the
 * maximum is taken on hi + lo, the winning child's pair is copied) */
int vxm_planar_to_ndhwc8_split_bf16(const float* const* planes, const long long* bstrides, int nplanes, void* out_hi,
                                    void* out_lo, int B, size_t V, void* stream);
int vxm_pool2_split_ndhwc_bf16(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int B, int Dc, int Hc, int Wc,
                               int C, int nd, void* stream);
/* out[c] = sum_{b,v} x[b][c][v] for planar fp32 x (B,C,V), C <= 32; work: 128*C floats */
int vxm_planar_channel_sums(const float* x, float* out, void* work, int B, int C, size_t V, void* stream);

/* ---- MaxPool(2) / nearest Upsample(2) + concat: reference networks.py:83-85,130,137-138 ----
 * pool factor is 2 on H, W and on D when D > 1 (nd == 3).  idx (uint8, same shape as y) stores the
 * argmax within the window for the backward. */
int vxm_maxpool2_fwd(const float* x, float* y, uint8_t* idx, int B, int C, int D, int H, int W,
                     int nd, void* stream);
int vxm_maxpool2_bwd(const float* grad_y, const uint8_t* idx, float* grad_x, int B, int C, int D,
                     int H, int W, int nd, void* stream);
/* out (B, Ca+Cb, 2D,2H,2W) = cat(upsample2(a (B,Ca,D,H,W)), skip (B,Cb,2D,2H,2W)); skip may be NULL (Cb=0) */
int vxm_upcat_fwd(const float* a, const float* skip, float* out, int B, int Ca, int Cb, int D,
                  int H, int W, int nd, void* stream);
/* grad_a overwritten (sum over the 2^nd children), grad_skip overwritten (may be NULL) */
int vxm_upcat_bwd(const float* grad_out, float* grad_a, float* grad_skip, int B, int Ca, int Cb,
                  int D, int H, int W, int nd, void* stream);

/* ---- Adam on one flat buffer: reference scripts/torch/train.py:161,220 (torch.optim.Adam) ----
 * p, g, m, v: n floats.  grad_scale multiplies g first (1/world_size after an allreduce-sum). */
int vxm_adam_step(float* p, const float* g, float* m, float* v, size_t n, int step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                  void* stream);

/* Same update with the step count kept on the device (incremented by the call): safe to capture in a CUDA graph. */
int vxm_adam_step_dev(float* p, const float* g, float* m, float* v, size_t n, int* step_counter, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VXM_B200_H */
