/* bf_b200.h -- C ABI of libbf_b200.so, the B200 (sm_100a) back-end for the
 * brainforge layer forward/backward + optimizer hot path.
 *
 * This is the drop-in boundary: every entry point replaces one operator of the
 * reference's `atomic` / `llatomic` op layer, its cost/metric functions or its
 * optimizer step (reference paths are relative to /root/reference/brainforge).
 * Plain pointers and sizes only; no torch / numpy types.  All `float*` /
 * `uint8_t*` arguments are DEVICE pointers unless the name starts with `h_`.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 * Tensors are dense, C-contiguous ("row-major"), FP32 in HBM; the reference
 * computes in float64 (config/numeric.py:1) -- see DESIGN.md for tolerances.
 *
 * Error convention: every function returns 0 on success, non-zero on failure;
 * bf_last_error() then returns a thread-local, NUL-terminated description.  The
 * Python shim maps BF_ERR_VALUE -> ValueError (tensor_op.py:18-21) and the rest
 * -> RuntimeError.  Nothing here ever falls back to a CPU implementation.
 */
#ifndef BF_B200_H
#define BF_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BF_API __attribute__((visibility("default")))
#else
#define BF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_ERR_CUDA 1     /* CUDA runtime / launch failure                  */
#define BF_ERR_VALUE 2    /* incompatible shapes (reference: ValueError)    */
#define BF_ERR_ARG 3      /* bad enum / null pointer / unsupported size     */
#define BF_ERR_WORKSPACE 4 /* bf_set_workspace() buffer too small           */

/* activation kinds: atomic/activation_op.py:157-159 registry */
enum { BF_ACT_LINEAR = 0, BF_ACT_SIGMOID = 1, BF_ACT_TANH = 2, BF_ACT_RELU = 3, BF_ACT_SOFTMAX = 4 };
/* cost kinds: metrics/costs.py:24-80 (mse, cxent, bxent, hinge) */
enum { BF_COST_MSE = 0, BF_COST_CXENT = 1, BF_COST_BXENT = 2, BF_COST_HINGE = 3 };
/* optimizer kinds: optimizers/__init__.py:4-6 */
enum { BF_OPT_SGD = 0, BF_OPT_MOMENTUM = 1, BF_OPT_NESTEROV = 2, BF_OPT_ADAGRAD = 3, BF_OPT_RMSPROP = 4, BF_OPT_ADAM = 5 };
/* conv modes: atomic/tensor_op.py:42-46 */
enum { BF_CONV_VALID = 0, BF_CONV_FULL = 1 };
/* arithmetic path of the contraction kernels */
enum {
  BF_PREC_FP32 = 0, /* FP32 FMA accumulate on the SIMT pipe: <=1e-5 rel. vs the float64 reference */
  BF_PREC_TF32 = 1  /* tcgen05 kind::tf32 tensor cores, FP32 accumulate in TMEM: looser bound */
};

/* ---- library / plumbing ------------------------------------------------ */
BF_API const char* bf_last_error(void);
BF_API int bf_version(void);
/* SM count and compute capability of the current device. */
BF_API int bf_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Scratch HBM for split-K partials etc.; owned by the caller (the Python host hands a torch buffer). */
BF_API int bf_set_workspace(void* d_ptr, size_t bytes);
/* Which kernel family the last contraction call on this thread used: 0 = SIMT FP32, 1 = tcgen05 TF32. */
BF_API int bf_last_path(void);
/* Number of kernels this library has launched so far in this process (for bench.py gpu_launches). */
BF_API unsigned long long bf_launch_count(void);
/* Upload of a PAGEABLE host array (the reference API's float64 NumPy inputs, layers' parameters, data sets): host threads
 * narrow float64 -> float32 (round to nearest even, as bf_cast_f64_to_f32 does on the device) or copy float32 into two
 * pinned staging buffers chunk by chunk while the previous chunk crosses the host link.  The source has been read when
 * the call returns; the device copies are ordered on `stream`.  (No reference counterpart: the reference has no device.) */
BF_API int bf_upload_host(float* d_dst, const void* h_src, size_t n, int src_is_f64, void* stream);
BF_API int bf_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
BF_API int bf_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
BF_API int bf_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
BF_API int bf_stream_sync(void* stream);
/* float64 <-> float32 on the device (host API speaks float64 like the reference, HBM holds FP32). */
BF_API int bf_cast_f64_to_f32(const double* d_src, float* d_dst, size_t n, void* stream);
BF_API int bf_cast_f32_to_f64(const float* d_src, double* d_dst, size_t n, void* stream);
/* (d0,d1,d2) -> (d1,d0,d2): layers/recurrent.py:26 `X.transpose(1, 0, 2)`. */
BF_API int bf_transpose01(const float* src, float* dst, int d0, int d1, int d2, void* stream);
/* dst[i] = a*x[i] + b  (NeuroEvolution.as_weights, learner/neuroevolution.py:20-22, is a=20,b=-10). */
BF_API int bf_affine(const float* x, float* dst, size_t n, float a, float b, void* stream);

/* ---- activations: atomic/activation_op.py ------------------------------- */
/* A = act(Z), elementwise; replaces Sigmoid/Tanh/ReLU/Linear.forward (:30,:54,:76,:87) and
 * llatomic/_llactivation.py:15-52.  In place allowed. */
BF_API int bf_act_forward(int kind, const float* Z, float* A, size_t n, void* stream);
/* Row softmax with max subtraction over axis 1: SoftMax.t1 (:127-130). In place allowed. */
BF_API int bf_softmax_forward(const float* Z, float* A, int m, int n, void* stream);
/* dout = din * act'(A) with the derivative expressed from the OUTPUT A (:33,:57,:79,:90,:134);
 * this is `delta *= activation.backward(output)` of layers/core.py:35,:52.  In place allowed. */
BF_API int bf_act_backward(int kind, const float* A, const float* din, float* dout, size_t n, void* stream);

/* out[r, i] = x[r, i] * (m ? m[i] : 1) * scale: DropOut (layers/fancy.py:60-90), whose mask has the shape of one
 * sample and is shared by the whole batch. */
BF_API int bf_mul_rowmask(const float* x, const float* m, float* out, size_t rows, size_t cols, float scale, void* stream);

/* ---- dense: atomic/core_op.py:9-20, llatomic/_llops.py:7-9 -------------- */
/* out = act(X@W + b); X (m,k), W (k,n), b (n) or NULL, out (m,n).  act fuses the layer's
 * activation (layers/core.py:29-31); BF_ACT_SOFTMAX applies the row softmax. */
BF_API int bf_dense_forward(const float* X, const float* W, const float* b, float* out, int m, int k, int n, int act,
                     int precision, void* stream);
/* DenseOp.backward(X,E,W) -> (dW,db,dX): dW = X^T@E (k,n), db = E.sum(0) (n), dX = E@W^T (m,k).
 * accumulate!=0 adds into dW/db (layers/core.py:37-38 `nabla_w += ...`), else assigns.
 * dX may be NULL (skipped). */
BF_API int bf_dense_backward(const float* X, const float* E, const float* W, float* dW, float* db, float* dX, int m, int k,
                      int n, int accumulate, int precision, void* stream);

/* 1 when, on BF_PREC_TF32, the dense contraction `site` (0 forward, 1 weight gradient, 2 data gradient) of this shape runs
 * on the TMA-staged tcgen05 GEMMs (operands read in place; the tensor core truncates them to TF32) rather than on the
 * gather-staged core (operands rounded to nearest); 0 for n <= 32 (FP32 FMA kernels).  atomic/core_op.py:12-20. */
BF_API int bf_dense_tma_supported(int m, int k, int n, int site);

/* Dense layer (n <= 32 outputs) directly behind a channels-last convolutional block: X / dX (m, hw*chan) are in
 * physical channels-last feature order (yx, c); W / dW (chan*hw, n) keep the reference's NCHW-flatten row order
 * (c, yx) (layers/core.py:82-86 Flatten + :27-39 Dense).  Saves the two layout-change passes around Flatten. */
/* 1 when bf_dense_forward_cl / bf_dense_backward_cl take a (k = chan*hw, n) layer (W fits the skinny kernels' shared
 * memory); callers fall back to a layout change + bf_dense_forward / bf_dense_backward otherwise. */
BF_API int bf_dense_cl_supported(int k, int n);
BF_API int bf_dense_forward_cl(const float* X, const float* W, const float* b, float* out, int m, int chan, int hw, int n, int act,
                        void* stream);
BF_API int bf_dense_backward_cl(const float* X, const float* E, const float* W, float* dW, float* db, float* dX, int m, int chan,
                         int hw, int n, int accumulate, void* stream);

/* Fused output head for learn_batch: the LAST Dense layer (n <= 32 outputs) forward, cost value, accuracy count, cost
 * derivative and the layer's backward pass in one kernel (+ one split reduction) -- layers/core.py:27-39,
 * metrics/costs.py:19-37, metrics/metrics.py:15-16, learner/backpropagation.py:19-22,26-29.
 *   out = act(X W + b);  cost_result[0] = cost(out, Y) (sum, not mean);  hits_result[0] = #(argmax out == argmax Y)
 *   delta = (out - Y) [* sample_w[row]] * act'(out);  dW (+)= X^T delta;  db (+)= delta.sum(0);  dX = delta W^T
 * X (m, chan*hw): with channels_last != 0 its columns are in physical (yx, c) order while W's rows keep the
 * reference's (c, yx) order (as bf_dense_forward_cl); otherwise chan*hw is simply k.  sample_w, hits_result, dX may be
 * NULL.  BF_COST_HINGE is not supported (its derivative is not out - Y). */
BF_API int bf_dense_head_supported(int k, int n, int act, int cost);
/* Whole learn_batch device step (minus the optimizer) of a two-layer perceptron at small batch in ONE kernel -- BASELINE
 * configs[0], 784-60-10 at batch 20 (layers/core.py:27-39 twice, metrics/costs.py:19-37, metrics/metrics.py:15-16,
 * learner/backpropagation.py:16-30): H = act1(X W1 + b1) (m,h); out = act2(H W2 + b2) (m,n); cost_result[0], hits_result[0]
 * as bf_dense_head; gW2 += H^T delta2, gb2 += delta2.sum(0), gW1 += X^T dH, gb1 += dH.sum(0) with
 * dH = (delta2 W2^T) * act1'(H).  A cluster of 8 CTAs splits the first layer's contraction index and exchanges the partial
 * sums through distributed shared memory; FP32 FMA in a fixed order (deterministic).  Limits: m <= 32, h <= 64, n <= 16,
 * k <= 1024; act1 in linear/sigmoid/tanh/relu; cost mse/cxent/bxent.  sample_w, hits_result may be NULL. */
BF_API int bf_mlp2_step_supported(int m, int k, int h, int n, int act1, int act2, int cost);
BF_API int bf_mlp2_step(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, const float* Y,
                 const float* sample_w, float* H, float* out, float* gW1, float* gb1, float* gW2, float* gb2, float* cost_result,
                 float* hits_result, int m, int k, int h, int n, int act1, int act2, int cost, void* stream);
BF_API int bf_dense_head(const float* X, const float* W, const float* b, const float* Y, const float* sample_w, float* out,
                  float* dW, float* db, float* dX, float* cost_result, float* hits_result, int m, int chan, int hw, int n,
                  int act, int cost, int channels_last, int accumulate, void* stream);

/* ---- convolution: atomic/tensor_op.py:6-75, llatomic/lltensor_op.py:9-40,75-121 ---- */
/* Y = act(corr(A,F)) + bias   (bias AFTER the activation: layers/tensor.py:77-78).
 * A (im,ic,iy,ix) NCHW, F (nf,ic,fy,fx), bias (nf) or NULL, stride 1.
 * VALID: Y (im,nf,iy-fy+1,ix-fx+1); FULL: zero padding f-1, Y (im,nf,iy+fy-1,ix+fx-1).
 * Returns BF_ERR_VALUE when filter depth != input depth (tensor_op.py:18-21). */
BF_API int bf_conv_forward(const float* A, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix,
                    int nf, int fc, int fy, int fx, int mode, int act, int precision, void* stream);
/* ConvolutionOp.backward(X,E,F) -> (dF,db,dX) for the VALID forward (tensor_op.py:49-61):
 * dF (nf,ic,fy,fx), db (nf) = E.sum((0,2,3)), dX (im,ic,iy,ix) = full(E, flip(F)^T).
 * Assigns (layers/tensor.py:83).  dF (with db) or dX may be NULL to skip that half. */
BF_API int bf_conv_backward(const float* X, const float* E, const float* F, float* dF, float* db, float* dX, int im, int ic,
                     int iy, int ix, int nf, int fy, int fx, int precision, void* stream);

/* ---- channels-last (NHWC) convolution on TMA + tcgen05 (the BF_PREC_TF32 fast path of ConvLayer) ----
 * Same arithmetic as bf_conv_forward / bf_conv_backward (atomic/tensor_op.py:9-61) on activations stored
 * (im, y, x, channel); the filter F stays (nf,ic,fy,fx) as the reference holds it.  Operands are staged by TMA
 * only (no im2col), a filter tap is a row offset into the staged tile.  Shapes: ic and nf multiples of 32
 * (bf_conv_nhwc_supported); everything else goes through bf_conv_forward / bf_conv_backward. */
BF_API int bf_conv_nhwc_supported(int ic, int nf, int fy, int fx);
BF_API int bf_conv_nhwc_forward(const float* X, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix,
                         int nf, int fy, int fx, int act, void* stream);
BF_API int bf_conv_nhwc_backward_data(const float* E, const float* F, float* dX, int im, int ic, int iy, int ix, int nf,
                               int fy, int fx, void* stream);
/* Pre-packed filter operands: bf_conv_nhwc_repack writes, in one launch, the forward operand Wt_fwd[tap][f][c] and the
 * flipped / transposed data-gradient operand Wt_dx[tap][c][f] (nf*ic*fy*fx floats each) of F (nf,ic,fy,fx); the two
 * _packed calls are bf_conv_nhwc_forward / bf_conv_nhwc_backward_data without their own re-layout launch.  A layer
 * repacks once per weight change (layers/tensor.py ConvLayer). */
BF_API int bf_conv_nhwc_repack(const float* F, float* Wt_fwd, float* Wt_dx, int nf, int ic, int fy, int fx, void* stream);
BF_API int bf_conv_nhwc_forward_packed(const float* X, const float* Wt_fwd, const float* bias, float* Y, int im, int ic, int iy, int ix,
                                int nf, int fy, int fx, int act, void* stream);
/* ConvLayer -> [Activation("relu")] -> PoolLayer(2) in ONE kernel (layers/tensor.py:28-30,75-79; atomic/tensor_op.py:93-107):
 * P (im, oy/2, ox/2, nf) = maxpool2x2([relu](corr(X, F) + bias)), mask = one byte per pooled element (bit dy*2+dx set
 * where the window element equals its max: multi-hot on ties, as bf_pool_nhwc_forward writes it).  The full-resolution
 * convolution output never reaches HBM.  BF_ERR_ARG when the plane has no whole-image tiling (caller runs the two
 * kernels). */
BF_API int bf_conv_nhwc_forward_pool_packed(const float* X, const float* Wt_fwd, const float* bias, float* P, uint8_t* mask, int im,
                                     int ic, int iy, int ix, int nf, int fy, int fx, int relu_in, void* stream);
BF_API int bf_conv_nhwc_backward_data_packed(const float* E, const float* Wt_dx, float* dX, int im, int ic, int iy, int ix, int nf,
                                      int fy, int fx, void* stream);
BF_API int bf_conv_nhwc_backward_filter(const float* X, const float* E, float* dF, float* db, int im, int ic, int iy,
                                 int ix, int nf, int fy, int fx, void* stream);
/* The same with db = sum of the `db_rows` rows of per-CTA channel sums the producer of E left in db_partial
 * (bf_pool_nhwc_backward_db): the kernel then spends no MMA on the bias gradient. */
BF_API int bf_conv_nhwc_backward_filter_dbp(const float* X, const float* E, float* dF, float* db, const float* db_partial, int db_rows,
                                     int im, int ic, int iy, int ix, int nf, int fy, int fx, void* stream);

/* First layer (ic <= 4 input channels, e.g. RGB): X stays NCHW as the reference holds it, the 32/64/128-channel
 * side (Y, E) is channels-last.  Forward = Toeplitz-descriptor implicit GEMM (no im2col), filter gradient = TMA-fed
 * E operand + SIMT-gathered patch rows; both on tcgen05 (atomic/tensor_op.py:9-32,49-55). */
BF_API int bf_conv_c3_supported(int ic, int iy, int ix, int nf, int fy, int fx);
BF_API int bf_conv_c3_forward(const float* X, const float* F, const float* bias, float* Y, int im, int ic, int iy, int ix,
                       int nf, int fy, int fx, int act, void* stream);
/* Convolution + [ReLU] + 2x2 max pooling in one pass: P (im,oy/2,ox/2,nf) and the tie mask (one byte per pooled
 * element, bits as in bf_pool_nhwc_forward) = MaxPoolOp.forward( [relu]( act(corr(X,F)) + bias ), 2 ), i.e.
 * ConvLayer -> [Activation("relu")] -> PoolLayer(2) (layers/tensor.py:28-30,75-79; layers/core.py:47-49).  The
 * full-resolution convolution output is never written to HBM (the host re-runs bf_conv_c3_forward if a caller reads
 * ConvLayer.output).  Needs an even output plane. */
BF_API int bf_conv_c3_forward_pool(const float* X, const float* F, const float* bias, float* P, uint8_t* mask, int im, int ic,
                            int iy, int ix, int nf, int fy, int fx, int act, int relu_in, void* stream);
BF_API int bf_conv_c3_backward_filter(const float* X, const float* E, float* dF, float* db, int im, int ic, int iy, int ix,
                               int nf, int fy, int fx, void* stream);

/* Filter gradient of ConvLayer -> [Activation("relu")] -> PoolLayer(2) straight from the pooling layer's delta: E =
 * MaxPoolOp.backward(dP, mask) (atomic/tensor_op.py:110-119) [* (gate > 0), layers/core.py:51-52] is formed on the fly
 * inside the kernel, the full-resolution delta never exists in HBM.  dP, gate (pooled forward output, or NULL for no
 * ReLU) (im,oy/2,ox/2,nf) channels-last, mask as written by bf_conv_c3_forward_pool / bf_pool_nhwc_forward. */
BF_API int bf_conv_c3_backward_filter_unpool(const float* X, const float* dP, const uint8_t* mask, const float* gate, float* dF,
                                      float* db, int im, int ic, int iy, int ix, int nf, int fy, int fx, void* stream);

/* ---- max pooling: atomic/tensor_op.py:78-128, llatomic/lltensor_op.py:43-72 ---- */
/* Bytes per window of the compact tie mask: 1 when fdim*fdim<=8, else 8 (fdim<=8). */
BF_API size_t bf_pool_mask_bytes(int im, int ic, int iy, int ix, int fdim);
/* MaxPoolOp.forward(A,fdim) -> (out, filt).  The multi-hot equality mask (bit dy*fdim+dx set for
 * EVERY element equal to the window max) is stored compactly, one word per window. */
BF_API int bf_pool_forward(const float* A, float* out, uint8_t* mask, int im, int ic, int iy, int ix, int fdim, void* stream);
/* MaxPoolOp.backward(E,filt) -> dX (im,ic,iy,ix) = filt * upsample(E).  Does not consume the mask. */
BF_API int bf_pool_backward(const float* E, const uint8_t* mask, float* dX, int im, int ic, int iy, int ix, int fdim,
                     void* stream);
/* Expands the compact mask to the reference's float `filt` array (A's shape, 0.0/1.0). */
BF_API int bf_pool_mask_expand(const uint8_t* mask, float* filt, int im, int ic, int iy, int ix, int fdim, void* stream);

/* ---- channels-last companions of the TMA convolution path (csrc/nhwc_ops.cu) ---- */
/* Layout change of an activation tensor: (im, c, hw) <-> (im, hw, c).  The reference holds every tensor NCHW
 * (atomic/tensor_op.py:9-15); the tcgen05 convolutions read and write channels-last, the host converts at the
 * edges of a convolutional block (Flatten, user reads of layer.output). */
BF_API int bf_layout_nchw_to_nhwc(const float* src, float* dst, int im, int c, int hw, void* stream);
BF_API int bf_layout_nhwc_to_nchw(const float* src, float* dst, int im, int c, int hw, void* stream);
/* MaxPoolOp.forward / backward (atomic/tensor_op.py:93-119) on channels-last tensors: A (im,iy,ix,c),
 * out / mask / E (im,oy,ox,c); mask words as in bf_pool_forward (same bits, channels-last order); c % 4 == 0.
 * relu_in != 0 applies max(0,.) to A while reading it: the Activation("relu") layer that precedes the pooling
 * layer (layers/core.py:50) costs no pass of its own.  gate != NULL (the pooled OUTPUT of the forward call)
 * multiplies E by (gate > 0): that layer's `delta * relu'(output)` (layers/core.py:52), exact because every
 * masked element of a window equals the pooled output. */
BF_API int bf_pool_nhwc_forward(const float* A, float* out, uint8_t* mask, int im, int c, int iy, int ix, int fdim, int relu_in,
                         void* stream);
BF_API int bf_pool_nhwc_backward(const float* E, const uint8_t* mask, const float* gate, float* dX, int im, int c, int iy,
                          int ix, int fdim, void* stream);
/* bf_pool_nhwc_backward that also writes, per CTA, the per-channel sums of dX into db_partial (bf_pool_nhwc_db_rows(..)
 * rows of c floats): summed over the rows this is E.sum((0,2,3)), the bias gradient of the convolution whose output
 * was pooled (atomic/tensor_op.py:55), handed to bf_conv_nhwc_backward_filter_dbp.  fdim 2, c dividing 1024. */
BF_API int bf_pool_nhwc_db_rows(int im, int c, int iy, int ix, int fdim);
BF_API int bf_pool_nhwc_backward_db(const float* E, const uint8_t* mask, const float* gate, float* dX, float* db_partial, int im, int c,
                             int iy, int ix, int fdim, void* stream);

/* ---- global average pooling: layers/tensor.py:95-114 ---- */
/* out (im, c) = A.mean over the hw pixels of every (image, channel) plane; nhwc != 0: A is stored (im, hw, c). */
BF_API int bf_gap_forward(const float* A, float* out, int im, int c, int hw, int nhwc, void* stream);
/* dX = repeat(delta / hw) over the plane (layers/tensor.py:106-110), written in the layout `nhwc` selects. */
BF_API int bf_gap_backward(const float* delta, float* dX, int im, int c, int hw, int nhwc, void* stream);

/* ---- costs / metrics: metrics/costs.py:11-37, metrics/metrics.py:13-16 ---- */
/* delta = (out - tgt) * (w ? w[row] : 1): CostFunction.derivative + learner/backpropagation.py:19-21. */
BF_API int bf_cost_delta(const float* out, const float* tgt, const float* w, float* delta, int m, int n, void* stream);
/* result[0] = cost: MSE = 0.5*sum((o-t)^2), CXENT = -sum(t*log o) (terms with t==0 contribute 0),
 * BXENT = -sum(t*log o + (1-t)*log(1-o)) (costs.py:48-49), HINGE = sum(max(0, 1 - t*o)) (costs.py:54-55). */
BF_API int bf_cost_value(int kind, const float* out, const float* tgt, int m, int n, float* result, void* stream);
/* Both of the above in one pass over the predictions; the delta is `out - tgt` for every kind except HINGE, whose
 * sub-derivative is -tgt where out <= 1 and 0 elsewhere (costs.py:58-66) (the training step needs the delta and the cost of the same
 * forward pass, learner/backpropagation.py:19-27). */
BF_API int bf_cost_delta_value(int kind, const float* out, const float* tgt, const float* w, float* delta, float* result, int m,
                        int n, void* stream);
/* result[0] = count(argmax_1(out) == argmax_1(tgt)), first maximum wins like numpy. */
BF_API int bf_accuracy(const float* out, const float* tgt, int m, int n, float* result, void* stream);

/* ---- LSTM: atomic/recurrent_op.py:46-112, llatomic/_lllstm.py:6-67 ---- */
/* LSTMOp(act).forward(X,W,b) -> (O,Z,cache).  X (T,B,D) time-major, W (D+H,4H) gate order
 * [cand|f|i|o], b (4H); O (T,B,H), Z (T,B,D+H), cache (6,T,B,H) = C,Ca,cand,f,i,o.
 * bw (5,T,B,H) or NULL: the five per-element coefficients of the backward recurrence (recurrent_op.py:94-104),
 * o*act'(C), i*act'(cand), C[t-1]*f', cand*i', Ca*o', with the derivatives (the reference's bwCa/bwgates, :85-88)
 * evaluated from the pre-activations, so that the saturated gates the +7 bias init produces do not lose their
 * derivative to FP32 cancellation in A(1-A). */
BF_API int bf_lstm_forward(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                    int T, int B, int D, int H, int act, int precision, void* stream);
/* Same with X batch-major (B,T,D) as the layer receives it: folds `X.transpose(1, 0, 2)` (layers/recurrent.py:26) into
 * the pass that builds Z. */
BF_API int bf_lstm_forward_bm(const float* X, const float* W, const float* b, float* O, float* Z, float* cache, float* bw,
                       int T, int B, int D, int H, int act, int precision, void* stream);
/* LSTMOp.backward(Z,O,E,W,cache) -> (dX (T,B,D), nablaW (D+H,4H), nablab (4H)); E (T,B,H) is read only.
 * bw as produced by bf_lstm_forward, or NULL to form the coefficients from the cached outputs in FP32.
 * dX may be NULL when the caller discards the input delta (learner/backpropagation.py:22 ignores it for the first layer).
 * Scratch (dgates, deltaC, split-K partials) comes from the workspace.  TF32 path with bw: the whole BPTT recurrence
 * is one persistent kernel (W_h slice resident in shared memory, deltaC in registers, per-batch-tile flags), dX and
 * the weight gradients are three large GEMMs over all T*B rows afterwards. */
BF_API int bf_lstm_backward(const float* Z, const float* O, const float* E, const float* W, const float* cache,
                     const float* bw, float* dX, float* nablaW, float* nablab, int T, int B, int D, int H, int act,
                     int precision, void* stream);

/* ---- plain recurrent layer: atomic/recurrent_op.py:9-43 (RLayer, layers/recurrent.py:50-77) ---- */
/* RecurrentOp(act).forward(X,W,b) -> (O,Z).  X (T,B,D) time-major, W (D+H,H), b (H); O (T,B,H), Z (T,B,D+H). */
BF_API int bf_rnn_forward(const float* X, const float* W, const float* b, float* O, float* Z, int T, int B, int D, int H,
                   int act, int precision, void* stream);
/* RecurrentOp.backward(Z,O,E,W) -> (dX, nablaW (D+H,H), nablab (H)).  E (T,B,H) is read only; Eout (T,B,H) receives
 * the array the reference leaves in E (each step scaled by act'(O[t]) and fed back through W[D:]); the reference's
 * dX is literally its slice [:, :, :D] (recurrent_op.py:40), which the caller takes. */
BF_API int bf_rnn_backward(const float* Z, const float* O, const float* E, const float* W, float* Eout, float* nablaW,
                    float* nablab, int T, int B, int D, int H, int act, int precision, void* stream);

/* ---- GRU layer: layers/recurrent.py:116-190 (implemented at layer level in the reference, no atomic op) ---- */
/* feedforward: X (T,B,D) time-major, W (D+H,3H) columns [U|R|O], b (3H) -> hs (T,B,H) hidden states, Z / Zk (T,B,D+H)
 * the operands [X[t], h] and [X[t], R*h] (self.Zs), gates (3,T,B,H) = U, R, O (self.gates). */
BF_API int bf_gru_forward(const float* X, const float* W, const float* b, float* hs, float* Z, float* Zk, float* gates, int T,
                   int B, int D, int H, int act, int precision, void* stream);
/* backpropagate: delta (T,B,H) -> dX (T,B,D), nablaW (D+H,3H), nablab (3H). */
BF_API int bf_gru_backward(const float* Z, const float* Zk, const float* gates, const float* delta, const float* W, float* dX,
                    float* nablaW, float* nablab, int T, int B, int D, int H, int act, int precision, void* stream);

/* ---- Clockwork layer: layers/recurrent.py:193-283 ---- */
/* feedforward: X (T,B,D), W (D+H,H), b (H), tick (H) = the layer's tick_array -> O (T,B,H) = self.cache,
 * Z (T,B,D+H) = self.Zs.  Column n updates at step t (1-based) iff t % tick[n] == 0, otherwise it sees a zero
 * pre-activation (recurrent.py:242-246). */
BF_API int bf_clockwork_forward(const float* X, const float* W, const float* b, const float* tick, float* O, float* Z, int T,
                         int B, int D, int H, int act, int precision, void* stream);
/* backpropagate: delta (T,B,H) -> dX (T,B,D), nablaW (D+H,H), nablab (H). */
BF_API int bf_clockwork_backward(const float* Z, const float* O, const float* delta, const float* W, const float* tick, float* dX,
                          float* nablaW, float* nablab, int T, int B, int D, int H, int act, int precision, void* stream);

/* ---- optimizer step on the flat vectors: optimizers/*.py, learner/backpropagation.py:37-42 ---- */
/* W <- optimize(W, g, m) in place.  s1 = velocity, s2 = memory (NULL where the rule has none).
 * hp: SGD{} MOMENTUM{mu} NESTEROV{mu} ADAGRAD{eps} RMSPROP{decay,eps} ADAM{decay_velocity,decay_memory,eps}.
 * zero_grad!=0 also clears g (zero_gradients, backpropagation.py:68-73) in the same pass. */
BF_API int bf_opt_step(int kind, float* W, float* g, float* s1, float* s2, size_t n, float eta, float m, float hp0, float hp1,
                float hp2, int zero_grad, void* stream);

/* Data-parallel form of the same step: all-reduce(sum) of the flat gradient vector over peer memory fused
 * with the optimizer update, PUSH form: every rank stores its vector into its slot of every peer's receive buffer
 * (NVLink stores), signals, waits for the world signals addressed to it, sums the world copies in rank order (identical
 * bits on every rank) and updates its replica -- one synchronisation per step.
 * G: this rank's gradient vector (ordinary device memory), n_total floats: [0, n_params) gradients, then a tail of step
 * scalars [cost sum, hits, sample count, -] that is summed, not optimised (the kernel posts m_local in slot 2).
 * recv / pads: DEVICE arrays of `world` pointers to every rank's receive buffer (2 * world * n_total floats, peer
 * mapped) and signal pad (blocks * world words, zero-initialised); epoch: zero-initialised local array of `blocks`
 * words; timeout: one word in mapped HOST memory, set to 1 if a peer did not arrive within ~10 s -- the CTAs that
 * gave up leave W and the optimizer state untouched and the caller must treat the step as failed.
 * Replaces `allreduce` + `optimizer.optimize` + `zero_gradients` (learner/backpropagation.py:37-42,68-73). */
BF_API int bf_opt_step_allreduce(int kind, float* W, float* G, float* s1, float* s2, size_t n_params, size_t n_total, float eta,
                          float m, float hp0, float hp1, float hp2, int zero_grad, void* const* recv, void* const* pads,
                          unsigned* epoch, unsigned* timeout, float m_local, int rank, int world, int blocks, void* stream);

/* x[0..n) = value without reading x (a NaN already there does not survive): zero_gradients / fresh biases
 * (learner/backpropagation.py:68-73, layers/abstract_layer.py:36-38). */
BF_API int bf_fill(float* x, size_t n, float value, void* stream);

/* ---- neuro-evolution fitness fan-out: evolution/_evolution.py:83-97 + learner/neuroevolution.py:20-22,34-37 ---- */
/* For each of P genomes (rows of `genomes`, `loci` = nin*nh+nh+nh*nout+nout floats in U(0,1)):
 * weights = (g-0.5)*20 -> Dense(nh,sigmoid) -> Dense(nout,softmax) on the shared X (N,nin) ->
 * fitness[p] = -sum(Y*log_softmax(logits))/N  (cxent/m, computed from log-softmax so FP32 does not
 * underflow where float64 does not).  Scratch for the hidden activations comes from the workspace. */
BF_API int bf_fitness_mlp(const float* genomes, const float* X, const float* Y, float* fitness, int P, int N, int nin, int nh,
                   int nout, int precision, void* stream);

/* ---- batched on-device gradient check: gradientcheck/raw_gradients.py:15-42 ---- */
/* The reference perturbs one weight at a time on the host and reads two costs back per parameter.  Here the whole
 * sweep stays on the device: a parameter counter lives in device memory, `bf_gradcheck_perturb` moves weight
 * W[offsets[*counter]] (phase 0: save it, set w+eps; phase 1: w-eps; phase 2: restore, ++*counter) and
 * `bf_gradcheck_cost` stores the double-precision cost of the current outputs in record[3*(*counter)+slot]
 * (slot 0: cost(w+eps), 1: cost(w-eps); element 2 holds the FP32 step actually taken, (w+eps)-(w-eps)).  One sweep
 * iteration (perturb, forward, cost, perturb, forward, cost, restore) is captured in a CUDA graph and replayed once
 * per parameter; the host reads `record` once at the end. */
BF_API int bf_gradcheck_perturb(float* W, const int* offsets, unsigned* counter, float* saved, double* record, float epsilon,
                         int phase, void* stream);
BF_API int bf_gradcheck_cost(int kind, const float* out, const float* tgt, int m, int n, double* record, const unsigned* counter,
                      int slot, void* stream);

/* ---- device-resident Population: evolution/_evolution.py:83-141,143-208,245-261 ---- */
/* The genome store G (Pstore, ldg) stays in HBM (ldg = loci rounded up to a multiple of 4 floats, padding zero) next to
 * Bt (Pstore*64, nin), the K-major TF32 first-layer operand of the fitness kernel:
 * Bt[g*64+h][k] = tf32((G[g][k*nh+h]-0.5)*20).  bf_population_fitness = bf_fitness_mlp for the `n` individuals listed
 * in `inds` (device int32, NULL: individuals 0..n-1), read in place: no genome upload, no re-layout pass
 * (`Population.update`, _evolution.py:83-97).  BF_PREC_TF32 only. */
BF_API int bf_population_fitness(const float* G, long ldg, int Pstore, const float* Bt, const int* inds, int n, const float* X,
                          const float* Y, float* fitness, int N, int nin, int nh, int nout, int precision, void* stream);
/* Refresh Bt for the listed genomes (NULL: genomes 0..n-1) after their rows of G changed. */
BF_API int bf_population_relayout(const float* G, long ldg, float* Bt, const int* inds, int n, int nin, int nh, void* stream);
/* One generation step in a single pass over the store (_evolution.py:181-183,245-261): for every individual i
 *   G_new[i] = survive[i] ? G_old[i] : where(coin < 0.5, G_old[left[i]], G_old[right[i]])      (mating)
 * then every locus is replaced by a fresh U(0,1) draw with probability `rate` (mutation); mutant[i] = 1 where a row
 * mutated.  Draws come from Philox4x32-10 keyed by (seed, generation), or -- when coin / mut_u / mut_val (P, loci) are
 * given -- from the caller (parity tests replay the reference's recorded draws).  G_new must not alias G_old. */
BF_API int bf_population_next_generation(const float* G_old, float* G_new, long ldg, int loci, int P, const unsigned char* survive,
                                  const int* left, const int* right, float rate, unsigned seed, unsigned generation,
                                  const float* coin, const float* mut_u, const float* mut_val, unsigned char* mutant,
                                  void* stream);
/* dst[r][:] = src[idx[r]][:] for rows of `row_floats` floats: the shuffled mini-batches of util/_nnutil.py:4-15 drawn from
 * a data set that was uploaded once (learner/abstract_learner.py:41-54 `fit`). */
BF_API int bf_gather_rows(const float* src, const int* idx, float* dst, int m, size_t row_floats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BF_B200_H */
