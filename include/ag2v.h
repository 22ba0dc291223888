/* libag2v_sm100a — C ABI of the B200-native AG2Vid hot path.
 *
 * The reference (roeiherz/AG2Video) has no FFI layer: its boundary is a set of
 * Python names bound at import time.  These entry points
 * are what a binding of that boundary calls; each one cites the reference code it
 * replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the
 *     caller (PyTorch's caching allocator in the shipped host code); the library
 *     never allocates or frees device memory and keeps no state between calls;
 *   - all work is enqueued on `stream` (a cudaStream_t); nothing synchronises;
 *   - return value 0 = ok, negative = error (ag2v_last_error_string() has the text;
 *     -1 bad argument, -2 CUDA error, -3 wrong architecture, -4 unsupported shape);
 *   - tensors are fp32; "NHWC" means channels-last storage of a logical NCHW tensor.
 */
#ifndef AG2V_H_
#define AG2V_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ag2v_stream_t; /* == cudaStream_t */

/* ---- library --------------------------------------------------------------- */
const char* ag2v_last_error_string(void);
int ag2v_version(void);
int ag2v_check_device(void);                 /* 0 iff the current device is sm_100 */
int ag2v_sm_count(void);
unsigned long long ag2v_launch_count(void);  /* kernels enqueued by this library so far */

/* ---- K1: action-graph convolution ---------------------------------------------
 * Replaces GraphTripleConv.forward (models/graph_models/graph.py:41-107) and its
 * autograd backward.  obj [B,O,Din], pred [B,E,Dp], edges [B,E,2] int64,
 * ind [B,E] uint8 (pred_indicators), net1 = (W1a [H,2Din+Dp], b1a, W1b [2H+Dpo,H], b1b),
 * net2 = (W2a [H,H], b2a, W2b [Dout,H], b2b); outputs new_obj [B,O,Dout], new_p [B,E,Dpo].
 * `saved` (ag2v_gcn_layer_saved_floats floats) carries activations to the backward.
 * All feature sizes must be multiples of 8. */
size_t ag2v_gcn_layer_saved_floats(int B, int O, int E, int H, int Dpo);
size_t ag2v_gcn_layer_bwd_workspace_floats(int B, int O, int E, int Din, int Dp, int H, int Dpo);
int ag2v_gcn_layer_fwd(const float* obj, const float* pred, const long long* edges, const uint8_t* ind,
                       const float* W1a, const float* b1a, const float* W1b, const float* b1b,
                       const float* W2a, const float* b2a, const float* W2b, const float* b2b,
                       int B, int O, int E, int Din, int Dp, int H, int Dout, int Dpo,
                       float* new_obj, float* new_p, float* saved, ag2v_stream_t stream);
int ag2v_gcn_layer_bwd(const float* obj, const float* pred, const long long* edges, const uint8_t* ind,
                       const float* W1a, const float* W1b, const float* W2a, const float* W2b,
                       const float* new_obj, const float* saved, const float* d_new_obj,
                       const float* d_new_p /* may be NULL */,
                       int B, int O, int E, int Din, int Dp, int H, int Dout, int Dpo,
                       float* workspace, float* dobj, float* dpred, float* dW1a, float* db1a, float* dW1b,
                       float* db1b, float* dW2a, float* db2a, float* dW2b, float* db2b, ag2v_stream_t stream);

/* ---- K1r: the Acts2LayoutModel recurrence as one persistent kernel per direction ---------------
 * Replaces the frame loop of Acts2LayoutModel.forward (models/graph_models/model.py:126-169):
 *   for t = 1..T-1: x = obj_vecs_net([emb | boxes[t-1]]); (x, p) = gconv_l(x, p, edges[t], ind[t]) for all layers
 *                   (GraphTripleConv.forward, graph.py:41-107); boxes[t] = boxes[t-1] + box_net(x)
 * One thread-block cluster of CS CTAs per chain = (model, clip); NC chains per launch.  The ten leading ints are
 * (O nodes <= 16, E edges per timestep <= 16, T frames, Kx = De = width of the attribute embedding, Dp, H, Dout, Dpo, NL).
 * Parameter order of `params` / `grads` (host arrays of device pointers): obj_vecs_net[0].weight [De][Kx+4],
 * obj_vecs_net[2].weight [De][De], per layer net1[0].{weight,bias}, net1[2].{weight,bias}, net2[0].{weight,bias},
 * net2[2].{weight,bias}, then box_net[0].{weight,bias}, box_net[2].{weight,bias}.
 * ag2v_recur_pack re-lays the parameters out per CTA (both directions) once per optimiser step. */
int ag2v_recur_cluster_size(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL);
int ag2v_recur_cluster_fits(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS);
int ag2v_recur_max_active_clusters(int CS);
int ag2v_recur_num_params(int NL);
int ag2v_recur_set_core(int core);      /* 0 = fp32 FFMA register tiles (default), 1 = 3xTF32 mma.sync; set before ag2v_recur_pack */
int ag2v_recur_get_core(void);
int ag2v_recur_set_profile(unsigned long long* buf);   /* debug timeline of CTA 0, see k1r_recur.cu */
size_t ag2v_recur_pack_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS);
size_t ag2v_recur_saved_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC);
size_t ag2v_recur_z_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC);
int ag2v_recur_pack(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS,
                    const float* const* params, float* pack, ag2v_stream_t stream);
/* emb [NC][O][Kx], pred [NC][T][E][Dp] (model-major, NC = n_models * clips); box0 [clips][O][4],
 * edges [clips][T][E][2] int64, ind [clips][T][E] uint8 are data shared by the models; objv [NC][T][O][Dout],
 * boxes [NC][T][O][4]; saved: ag2v_recur_saved_floats floats kept for the backward. */
int ag2v_recur_fwd(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS, int NC,
                   int n_models, const float* const* packs, const float* emb, const float* box0, const float* pred,
                   const long long* edges, const unsigned char* ind, float* objv, float* boxes, float* saved,
                   ag2v_stream_t stream);
/* Reverse chain of ONE model = chains [chain0, chain0 + nchains) of the forward launch: data gradients and the
 * Z buffer (gated output gradient of every Linear) for ag2v_recur_wgrad.  Model-local arrays: boxes, d_objv and
 * d_boxes (may be NULL = zero), d_emb, d_box0, d_pred, and the per-chain partial sums dw0box [nchains][De][4],
 * dwb2 [nchains][4][H], dbb2 [nchains][4]. */
int ag2v_recur_bwd(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS, int NC,
                   int chain0, int nchains, const float* pack, const float* saved, const float* boxes,
                   const float* d_objv, const float* d_boxes, const long long* edges, const unsigned char* ind,
                   float* z, float* d_emb, float* d_box0, float* d_pred, float* dw0box, float* dwb2, float* dbb2,
                   ag2v_stream_t stream);
/* All weight / bias gradients of the model as ONE grouped GEMM over the (chain, t, row) rows + one column-sum
 * kernel; obj_vecs_net[0].weight receives its first Kx columns, box_net[2] comes from dwb2 / dbb2. */
int ag2v_recur_wgrad(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC,
                     int chain0, int nchains, const float* saved, const float* z, const float* emb,
                     float* const* grads, ag2v_stream_t stream);

/* ---- K2: layout composition ---------------------------------------------------
 * Replaces boxes_to_layout (models/layout.py:28-63) incl. _boxes_to_grid (:98-130)
 * and _pool_samples (:205-237), batched over N (clip, frame) samples.
 * vecs [N,O,D], boxes [N,O,4] xywh, valid [N,O] uint8 or NULL (the callers' object
 * mask, models/utils.py:95-102), lin_x [W] / lin_y [H] = torch.linspace(0,1,n) as the
 * reference computes it (on the CPU), out [N,D,H,W].  All-zero boxes are dropped
 * (layout.py:40-42).  avg != 0 selects pooling='avg'.  The workspace filled by the
 * forward is reused by the backward (recompute = 0) or rebuilt from boxes. */
size_t ag2v_boxes_to_layout_workspace_bytes(int N, int O, int H, int W);
int ag2v_boxes_to_layout_fwd(const float* vecs, const float* boxes, const uint8_t* valid, const float* lin_x,
                             const float* lin_y, int N, int O, int D, int H, int W, int avg, void* workspace,
                             float* out, ag2v_stream_t stream);
int ag2v_boxes_to_layout_bwd(const float* dout, const float* boxes, const uint8_t* valid, const float* lin_x,
                             const float* lin_y, int N, int O, int D, int H, int W, int avg, int recompute,
                             void* workspace, float* dvecs, ag2v_stream_t stream);
/* Same kernels with a batch stride (in floats, >= D*H*W) on the image side: sample n lives at
 * out + n*out_batch_stride as [D,H,W].  Writes the layout straight into a channel slice of a
 * wider NCHW buffer - the discriminator's cat([img, seg], dim=channels)
 * (spade_models/networks/discriminator.py:338-342) - and reads its gradient from the same
 * slice, so neither the 256-channel layout nor the concatenation copy exists. */
int ag2v_boxes_to_layout_fwd_strided(const float* vecs, const float* boxes, const uint8_t* valid, const float* lin_x,
                                     const float* lin_y, int N, int O, int D, int H, int W, int avg, void* workspace,
                                     float* out, long long out_batch_stride, ag2v_stream_t stream);
int ag2v_boxes_to_layout_bwd_strided(const float* dout, long long dout_batch_stride, const float* boxes,
                                     const uint8_t* valid, const float* lin_x, const float* lin_y, int N, int O, int D,
                                     int H, int W, int avg, int recompute, void* workspace, float* dvecs,
                                     ag2v_stream_t stream);
/* dboxes [N,O,4]: gradient with respect to the xywh boxes - autograd of _boxes_to_grid and grid_sample's grid
 * gradient (layout.py:55-57, :119-128); only the border pixels of each object's window contribute.  `workspace` is the
 * buffer the forward filled; dropped objects get zero. */
int ag2v_boxes_to_layout_dboxes(const float* dout, long long dout_batch_stride, const float* vecs, const float* boxes,
                                const float* lin_x, const float* lin_y, int N, int O, int D, int H, int W,
                                void* workspace, float* dboxes, ag2v_stream_t stream);

/* ---- K4: layout fused into its consumer convolution ----------
 * conv3x3(layout)[co,p] = sum_o sum_k U[o,k,co] * m_o(p+k) with U[o,k,:] = W[:,:,k] v[o,:]: the
 * 1024-channel layout that generator.py:38-54 materialises and the dense 1027->512 / 1027->32
 * convolutions over it (generator.py:29-33,82-83; flows_generator.py:32) collapse to these
 * memory-bound kernels (U is a tiny GEMM done by the host).  tables = workspace of
 * ag2v_boxes_to_layout_workspace_bytes(N,S,H,W) filled by ag2v_boxes_to_layout_tables for
 * boxes [N,S,4]; out/dout are NHWC [N,H,W,Co], Co in {32, 512}; U/dU are [N,S,9,Co]. */
int ag2v_boxes_to_layout_tables(const float* boxes, const uint8_t* valid, const float* lin_x, const float* lin_y,
                                int N, int O, int H, int W, void* workspace, ag2v_stream_t stream);
size_t ag2v_layout_conv_bwd_workspace_floats(int N, int S, int Co, int H);
int ag2v_layout_conv_fwd(const float* U, const void* tables, int N, int S, int Co, int H, int W, float* out,
                         ag2v_stream_t stream);
int ag2v_layout_conv_bwd(const float* dout, const void* tables, int N, int S, int Co, int H, int W, float* part,
                         float* dU, ag2v_stream_t stream);

/* ---- K7: the discriminator's PatchGAN stem through the rank-1 layout ----
 * Replaces boxes_to_layout + torch.cat([img, seg]) + the dense 4x4 stride-2 convolution 259 -> 64 of
 * NLayerActionDiscriminator.model0 (spade_models/networks/discriminator.py:317-342, :326-372) and, for
 * the coarser scales, the avg_pool2d(3, stride 2, padding 1, count_include_pad=False) of the
 * concatenation (:271, :350):
 *   conv(seg)[co,p] = sum_o sum_k U[o,k,co] * m_o(stride*p + k - pad),  U[o,k,:] = W[:,3:,k] v[o,:]
 * tables / tables_out: workspaces of ag2v_boxes_to_layout_workspace_bytes at (H,W) / ((H-1)/2+1,(W-1)/2+1);
 * U, dU [N,S,kernel*kernel,Co]; out/dout NHWC [N,Ho,Wo,Co], Ho = (H + 2*pad - kernel)/stride + 1, updated
 * in place (the caller puts the convolution of the image channels + bias there first).
 * Built for kernel 4, stride 2, pad 2, Co = 64; anything else returns an argument error. */
int ag2v_layout_tables_avgpool(const void* tables, int N, int O, int H, int W, void* tables_out, ag2v_stream_t stream);
size_t ag2v_layout_sconv_bwd_workspace_floats(int N, int S, int Co, int Ho, int KK);
int ag2v_layout_sconv_fwd(const float* U, const void* tables, int N, int S, int Co, int H, int W, int kernel,
                          int stride, int pad, float* out, ag2v_stream_t stream);
int ag2v_layout_sconv_bwd(const float* dout, const void* tables, int N, int S, int Co, int H, int W, int kernel,
                          int stride, int pad, float* part, float* dU, ag2v_stream_t stream);

/* K5 — spectral normalisation of up to 48 convolution weights in ONE launch (the spectral_norm
 * wrappers of SPADEResnetBlock, architecture.py:34-41, and of get_nonspade_norm_layer,
 * normalization.py:16-50; arithmetic of torch/nn/utils/spectral_norm.py compute_weight).
 * All array arguments are HOST arrays of n entries; w/out/u/v/grad entries are device pointers.
 * Weight i is [co, cin, kh, kw] with taps = kh*kw, stored contiguous (channels_last[i] = 0) or
 * channels_last (1).  power_iteration != 0: v <- normalize(W^T u), u <- normalize(W v) written
 * back to u / v (training mode); else u / v are only read (eval mode).  out = W / (u . W v).
 * `save` (sizes from ag2v_spectral_norm_sizes) receives sigma, u and v for the backward:
 * grad_w = G / sigma - (<G,W> / sigma^2) u v^T.  Fixed-order reductions (deterministic). */
int ag2v_spectral_norm_sizes(int n, const int* co, const int* cin, const int* taps, size_t* save_floats,
                             size_t* fwd_scratch_floats, size_t* bwd_scratch_floats);
int ag2v_spectral_norm_fwd(int n, const void* const* w, void* const* out, void* const* u, void* const* v,
                           const int* co, const int* cin, const int* taps, const int* channels_last, float* save,
                           size_t save_floats, float* scratch, size_t scratch_floats, int power_iteration, float eps,
                           ag2v_stream_t stream);
int ag2v_spectral_norm_bwd(int n, const void* const* w, const void* const* grad_out, void* const* grad_w,
                           const int* co, const int* cin, const int* taps, const int* channels_last,
                           const float* save, size_t save_floats, float* scratch, size_t scratch_floats,
                           ag2v_stream_t stream);

/* K5 sigma mode, for a call that batches `iters` frame groups: `iters` successive power iterations,
 * each saving its sigma / u / v (save: iters * save_floats of ag2v_spectral_norm_sizes; it starts
 * with sigma [iters][(n+3)&~3]).  No weight is written: group g's convolution runs on weight_orig
 * and its output is scaled by 1/sigma_g.  Backward of the sigmas: grad_w[i] = sum_it
 * dsigma[it][i] u_it v_it^T (dsigma: device array [iters][n]). */
int ag2v_spectral_norm_sigma_fwd(int n, const void* const* w, void* const* u, void* const* v, const int* co,
                                 const int* cin, const int* taps, const int* channels_last, int iters, float* save,
                                 size_t save_floats, float* scratch, size_t scratch_floats, int power_iteration,
                                 float eps, ag2v_stream_t stream);
int ag2v_spectral_norm_sigma_bwd(int n, const float* dsigma, void* const* grad_w, const int* co, const int* cin,
                                 const int* taps, const int* channels_last, int iters, const float* save,
                                 size_t save_floats, ag2v_stream_t stream);
/* The same gradient for ONE weight (`only`) of the list, straight from the gradients of its 1/sigma scales: dscale_g
 * [iters] per frame group and / or dscale_img [iters * images_per_group] per image (either may be null); coefficients
 * -(dscale) / sigma^2.  One launch per weight from that weight's own autograd node, so weight gradients complete layer
 * by layer and the data-parallel gradient all-reduce overlaps the rest of the backward pass. */
int ag2v_spectral_norm_scale_bwd_one(int n, int only, const float* dscale_g, const float* dscale_img,
                                     int images_per_group, void* grad_w, const int* co, const int* cin, const int* taps,
                                     const int* channels_last, int iters, const float* save, size_t save_floats,
                                     ag2v_stream_t stream);

/* K6 — 3x3 convolutions with 2-3 output channels at full resolution (the generator's conv_img
 * 64 -> 3 between LeakyReLU(0.2) and tanh, spade_models/networks/generator.py; the flow head
 * conv_flow 32 -> 2, flows_generator.py): streaming kernels, y = act_out(conv(act_in(x)) + bias).
 * x [B,H,W,CI], y / dy [B,H,W,CO] NHWC; w [CO][CI][3][3] or channels-last [CO][3][3][CI]
 * (w_channels_last); slope_in = LeakyReLU slope in front (1 = none); act_out 0 none, 1 tanh.
 * Instantiated for (CI, CO) = (64, 3) and (32, 2). */
int ag2v_thin_conv3x3_supported(int CI, int CO);
size_t ag2v_thin_conv3x3_workspace_floats(int B, int H, int W, int CI, int CO);
int ag2v_thin_conv3x3_fwd(const float* x, const float* w, const float* bias, int B, int H, int W, int CI, int CO,
                          int w_channels_last, float slope_in, int act_out, float* y, ag2v_stream_t stream);
int ag2v_thin_conv3x3_bwd(const float* x, const float* w, const float* dy, const float* y, int B, int H, int W, int CI,
                          int CO, int w_channels_last, float slope_in, int act_out, float* dx, float* workspace,
                          float* dw, float* dbias, ag2v_stream_t stream);

/* masks_to_layout (models/layout.py:66-95, _pool_mask_samples :164-202) for one
 * (clip, frame): vecs [O,D], boxes [O,4] xywh, masks [O,M,M]; S [O,H,W] receives the
 * sampled masks (kept for the backward); test_mode != 0 composites objects in
 * ascending order of sampled mass (order = O ints of scratch); out [D,H,W]. */
int ag2v_masks_to_layout_fwd(const float* vecs, const float* boxes, const float* masks, const float* lin_x,
                             const float* lin_y, int O, int D, int M, int H, int W, int test_mode, float* S,
                             int* order, float* out, ag2v_stream_t stream);
int ag2v_masks_to_layout_bwd(const float* dout, const float* S, int O, int D, int H, int W, float* dvecs,
                             ag2v_stream_t stream);
/* Gradients of the same call (train mode) with respect to the masks [O,M,M] and the xywh boxes [O,4] - what autograd
 * gives the reference through F.grid_sample (layout.py:87-91); either output may be null.  G = scratch [O,H,W];
 * dmasks must be zero-initialised (atomics, like torch's grid_sampler backward). */
int ag2v_masks_to_layout_bwd_inputs(const float* dout, const float* vecs, const float* boxes, const float* masks,
                                    const float* lin_x, const float* lin_y, int O, int D, int M, int H, int W, float* G,
                                    float* dmasks, float* dboxes, ag2v_stream_t stream);

/* crop_bbox (models/bilinear.py:102-131 with tensor_linspace :192-221) over a flat
 * list of crops: feats [NF,C,H,W] NCHW, frame[n] = source image of crop n, boxes
 * [n,4] xywh, ws/we = torch.linspace(1,0,steps) / (0,1,steps) for WW (x) and HH (y);
 * out [n,C,HH,WW].  The backward adds into a zero-initialised dfeats (atomics). */
int ag2v_crop_bbox_fwd(const float* feats, const int* frame, const float* boxes, const float* ws_x,
                       const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C, int H, int W,
                       int HH, int WW, float* out, ag2v_stream_t stream);
int ag2v_crop_bbox_bwd(const float* dout, const int* frame, const float* boxes, const float* ws_x,
                       const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C, int H, int W,
                       int HH, int WW, float* dfeats, ag2v_stream_t stream);
/* crop_bbox(backend='jj') (models/bilinear.py:127-128 -> bilinear_sample :134-189): coordinates in [0,1] scaled by the
 * source size, taps clamped to the image, the reference's weight and summation order.  Arguments as above. */
int ag2v_crop_bbox_jj_fwd(const float* feats, const int* frame, const float* boxes, const float* ws_x,
                          const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C, int H, int W,
                          int HH, int WW, float* out, ag2v_stream_t stream);
int ag2v_crop_bbox_jj_bwd(const float* dout, const int* frame, const float* boxes, const float* ws_x,
                          const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C, int H, int W,
                          int HH, int WW, float* dfeats, ag2v_stream_t stream);

/* ---- K3: SPADE ---------------------------------------------------------------
 * Pieces of SPADE.forward (models/spade_models/networks/normalization.py:96-110),
 * the LeakyReLU of SPADEResnetBlock (architecture.py:53-54,68) and their backward.
 * Host code composes them (ag2video_b200/spade.py); the SyncBN all-reduce of the
 * per-channel sums (sync_batchnorm/batchnorm.py:74-83) sits between stats and finalize. */

/* Groups: the frames of a clip can be batched into ONE call while keeping the per-call batch
 * statistics of the reference's frame loop (generator.py:56-94 calls every layer once per frame):
 * every activation is then [groups][P][C] (group-major batch) and statistics are per group. */

/* per-channel sums over x [groups][P,C] NHWC: sums[g][0..C) = sum x, sums[g][C..2C) = sum x^2 (doubles);
 * partial: groups * ag2v_chan_partial_floats(P, C, NS) floats */
size_t ag2v_chan_partial_floats(long long P, int C, int NS);
int ag2v_bn_stats(const float* x, long long P, int C, int groups, float* partial, double* sums, ag2v_stream_t stream);
/* F.batch_norm(training) statistics (normalization.py:99): mean/rstd [groups][C] + running update
 * (once per group, in group order); count = elements per channel of one group; unbias_count
 * (0 = count) is the n of the running estimate's n/(n-1) factor: statistics of a nearest-2x
 * up-sampled tensor are taken on its low-resolution source with unbias_count = 4 * count. */
/* in_scale (optional, [groups]): the layer's input is in_scale[g] * x while the sums were taken on x
 * (a spectrally normalised convolution evaluated on weight_orig): BN(s x; eps) == BN(x; eps / s^2),
 * so the 1/sigma multiplication never touches the activation. */
int ag2v_bn_finalize(const double* sums, double count, double unbias_count, int C, int groups, float eps,
                     float momentum, const float* in_scale, float* running_mean, float* running_var, float* mean,
                     float* rstd, ag2v_stream_t stream);
/* y = act((x - mean_g) * rstd_g * weight + bias): the affine SyncBN + LeakyReLU(0.2) stages
 * (normalization.py:16-50); slope 1 = no activation.  Its backward is ag2v_spade_bwd_pre with
 * chan_gamma = 1 (gamma := weight [C], dgb = NULL) followed by ag2v_spade_bwd_dx. */
int ag2v_bn_act_fwd(const float* x, const float* mean, const float* rstd, const float* weight, const float* bias,
                    long long P, int C, int groups, float slope, float* y, ag2v_stream_t stream);
/* eval mode: mean / rstd [groups][C] from the running estimates (in_scale as above) */
int ag2v_bn_eval_stats(const float* running_mean, const float* running_var, int C, int groups, float eps,
                       const float* in_scale, float* mean, float* rstd, ag2v_stream_t stream);

/* OIHW 3x3 weights -> [9][Nout][Cin] (dgrad = 1: transposed + flipped for the input
 * gradient).  With wb != NULL, (wa, wb) = (mlp_gamma, mlp_beta) are interleaved in
 * groups of 8 channels so one GEMM yields gamma and beta in the same thread. */
int ag2v_pack_w3x3(const float* wa, const float* wb, const float* ba, const float* bb, int Co, int Ci, int dgrad,
                   int round_ops, float* dst, float* bias_dst, ag2v_stream_t stream);

/* 3x3, pad 1 implicit-GEMM convolution on an NHWC view (element strides in_s*; the
 * nearest down-sample of normalization.py:102 is a strided view of the segmap):
 *   epilogue 0: +bias   1: relu(+bias) (mlp_shared, :103)
 *            2: SPADE — Nout = 2C gamma|beta, out = act((x-mean)*rstd*(1+gamma)+beta) (:104-108)
 *            3: out = gate > 0 ? acc : 0     4: out += acc (gradient into the shared segmap)
 * impl 0 = auto, 1 = mma.sync kernel, 2 = tcgen05 kernel, 3 = mma.sync with 3xTF32 products
 * (fp32-class accuracy, validation mode).  round_ops / round_out: round the stored GEMM operands
 * to nearest TF32 (the tcgen05 TF32 path truncates).  group_pixels > 0 (epilogue 2): mean / rstd
 * are [groups][C] and pixel p uses group p / group_pixels.  x_up != 0 (epilogue 2): x is
 * [B, Hh/2, Ww/2, C] and is read through a nearest 2x up-sampling (the `up(x)` in front of every
 * SPADEResnetBlock of the generator, spade_models/networks/generator.py) that is never
 * materialised.  Epilogue 0 also takes scale [groups]
 * (out = acc * scale[group] + bias, the 1/sigma of a spectrally normalised convolution evaluated
 * on weight_orig, architecture.py:34-41) and res, a dense [P, Nout] tensor added to the result
 * (the residual sum x_s + dx of SPADEResnetBlock, architecture.py:62); both may be NULL. */
int ag2v_conv3x3(const float* in, long long in_sb, long long in_sy, long long in_sx, int B, int Hh, int Ww, int Cin,
                 const float* wpk, const float* bias, int Nout, float* out, long long out_sb, long long out_sy,
                 long long out_sx, int epilogue, int round_out, const float* x, const float* mean,
                 const float* rstd, float* gamma_out, float slope, int C, long long group_pixels, int x_up,
                 const float* scale, const float* res, const float* gate, float* splitk_ws, size_t splitk_ws_floats,
                 int impl, ag2v_stream_t stream);
/* split-K scratch (floats) that lets low-resolution layers use the whole chip; 0 = none needed */
size_t ag2v_conv3x3_splitk_floats(int B, int Hh, int Ww, int Cin, int Nout);
int ag2v_conv3x3_tc_supported(int B, int Hh, int Ww, int Cin, int Nout, int epilogue);
/* BN statistics fused into the producing convolution's epilogue (the SPADE / batch norm that consumes y needs
 * sum y and sum y^2 per channel and statistics group, normalization.py:99): warp-shuffle column sums per M tile in the
 * tcgen05 epilogue + a fixed-order reduction over tiles.  stat_part: mtiles * 2 * Nout floats of scratch. */
int ag2v_conv3x3_stats_info(int B, int Hh, int Ww, int Cin, int Nout, int groups, int* mtiles, int* tiles_per_group);
int ag2v_conv3x3_bias_stats(const float* in, int B, int Hh, int Ww, int Cin, const float* wpk, const float* bias,
                            int Nout, float* out, int round_out, long long group_pixels, const float* scale,
                            const float* res, int groups, float* stat_part, double* sums, ag2v_stream_t stream);

/* SPADE backward, element-wise: pass 1 produces d(gamma|beta) [P,2C], dxhat and the
 * per-channel sums [groups][5][C] (sum g, sum g*xhat, sum dxhat, sum dxhat*xhat, sum dout*out);
 * pass 2 turns dxhat into dx (batch-norm backward) in place.  The fifth sum, added over channels
 * and divided by the group's scale, is d(scale) of the scaled convolution that consumed `out`.  chan_gamma = 1: `gamma` is a per-channel scale
 * [C] (affine batch norm) and dgb may be NULL. */
int ag2v_spade_bwd_pre(const float* dout, const float* out, const float* x, const float* gamma, const float* mean,
                       const float* rstd, long long P, int C, int groups, int act, float slope, int round_ops,
                       int chan_gamma, int up_h, int up_w, float* dgb, float* dxhat, float* partial, double* sums,
                       ag2v_stream_t stream);
/* up_h / up_w > 0 (both calls): x is the low-resolution source [.., up_h/2, up_w/2, C] of a nearest
 * 2x up-sampling; P counts full-resolution pixels per group and pass 2 writes the gradient of the
 * low-resolution x (sum over the four children) to dx_low instead of working in place. */
int ag2v_spade_bwd_dx(const float* x, float* dxhat, const float* mean, const float* rstd, const double* sums,
                      double count, int training, long long P, int C, int groups, int up_h, int up_w, float* dx_low,
                      ag2v_stream_t stream);

/* weight gradient of a 3x3 conv: split-K partials (tcgen05 MN-major kernel or mma.sync,
 * `impl` as for ag2v_conv3x3), then reduction + scatter to OIHW */
int ag2v_wgrad3x3_nsplit(int B, int Hh, int Ww, int Nout, int Cin, int impl);
int ag2v_wgrad3x3(const float* dy, int Nout, const float* x, long long x_sb, long long x_sy, long long x_sx, int Cin,
                  int B, int Hh, int Ww, float* part, int impl, ag2v_stream_t stream);
int ag2v_unpack_dw3x3(const float* part, int nsplit, int Co, int Ci, int two, float* dwa, float* dwb,
                      ag2v_stream_t stream);
/* Channels-last ([Co][3][3][Ci]) variants of ag2v_pack_w3x3 / ag2v_unpack_dw3x3 (weights of a
 * channels_last module are packed without a layout copy; the dgrad form is a tiled transpose), and
 * the pre-pass over the output gradient of the main convolutions of SPADEResnetBlock
 * (architecture.py:34-41,56-62), which run on the same implicit-GEMM kernels. */
int ag2v_pack_w3x3_cl(const float* wa, const float* wb, const float* ba, const float* bb, int Co, int Ci, int dgrad,
                      int round_ops, float* dst, float* bias_dst, ag2v_stream_t stream);
int ag2v_unpack_dw3x3_cl(const float* part, int nsplit, int Co, int Ci, int two, float* dwa, float* dwb,
                         ag2v_stream_t stream);
/* dys = tf32(dy * scale[group]) on [groups][P][C]; sums[g][0..C) = sum dy (doubles) */
int ag2v_scaled_grad_pre(const float* dy, const float* scale, long long P, int C, int groups, float* dys,
                         float* partial, double* sums, ag2v_stream_t stream);
/* dst[i] = (float) sum over groups of src[g * group_stride + i] */
int ag2v_double_to_float(const double* src, int n, int groups, long long group_stride, float* dst,
                         ag2v_stream_t stream);
/* dst = round-to-nearest TF32 of src (same layout; the tcgen05 TF32 path truncates, so GEMM
 * operands are rounded where they are produced; this covers the externally produced segmap) */
int ag2v_round_tf32(const float* src, float* dst, long long n, ag2v_stream_t stream);

/* ---- K8: SyncBN statistics over NVLink peer memory ------------------------------------------------------------
 * Replaces the exchange of the per-channel sums between the replicas of SynchronizedBatchNorm
 * (models/spade_models/networks/sync_batchnorm/batchnorm.py:74-83 forward, :105-145 backward; there a master / slave
 * queue between device threads, comm.py).  One process per GPU: each rank allocates a window, sends its 64-byte CUDA IPC
 * handle to the others over any host channel and maps theirs; ag2v_peer_allreduce_f64 is then ONE kernel per rank that
 * stores the rank's vector into every window, publishes its call number and sums the `world` rows in rank order
 * (bit-identical totals on all ranks).  The call number is counted in the window on the device, so a captured launch
 * replays correctly; calls on one window set must be device-ordered and identical on all ranks.  cap even, <= 32768. */
size_t ag2v_peer_window_bytes(int world, int cap);
int ag2v_peer_window_alloc(int world, int cap, void** window);
int ag2v_peer_window_free(void* window);
int ag2v_peer_window_export(void* window, unsigned char* handle64);
int ag2v_peer_window_import(const unsigned char* handle64, void** window);
int ag2v_peer_window_close(void* window);
int ag2v_peer_allreduce_f64(double* vec, int n, void* const* windows, int rank, int world, int cap,
                            ag2v_stream_t stream);

/* K8b: gradient all-reduce(avg) with the copy engines doing the transport between CUDA-IPC windows (replaces the NCCL
 * all-reduce of the data-parallel step when selected; the reference's nn.DataParallel gathers gradients on device 0,
 * torch/nn/parallel).  ag2v_peer_alloc: exportable device memory (export / import / close / free as for the windows).
 * Per bucket: ag2v_ce_sync (post + wait on flags carrying the bucket's exchange count, kept on the device),
 * ag2v_peer_memcpy (cudaMemcpyAsync between mapped windows), ag2v_ce_reduce (own = (own + staged copies) * scale), see
 * csrc/k8_peer.cu.  Flag block = ag2v_ce_flag_bytes() bytes, zero-initialised, one per rank. */
int ag2v_peer_alloc(size_t bytes, void** ptr);
size_t ag2v_ce_flag_bytes(void);
int ag2v_peer_memcpy(void* dst, const void* src, size_t bytes, ag2v_stream_t stream);
int ag2v_ce_sync(void* const* flags, int rank, int world, int bucket, int phase, int bump, int post, int wait, int back,
                 ag2v_stream_t stream);
int ag2v_ce_reduce(float* own, const float* staged, int parts, long long n, float scale, int ctas, ag2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AG2V_H_ */
