// cg_net.cuh -- network weight table + trunk launch interface (internal).
#pragma once
#include "cg_common.cuh"

// One folded (conv|linear)+BN layer: Wt is [K][C] row-major (k-major), b is [C].
struct cg_layer {
  const float *Wt;
  const float *b;
  int K, C;
};

// Weight order inside the blob (must match catgrasp_b200/weights.py:BLOB_ORDER).
enum cg_layer_id {
  L_S3_C1 = 0, L_S3_C2, L_S3_C3, L_S3_F1, L_S3_F2, L_S3_F3,   // STN3d      pointnet2.py:153-186
  L_E_C1,                                                      // encoder conv1 :252
  L_SK_C1, L_SK_C2, L_SK_C3, L_SK_F1, L_SK_F2, L_SK_F3,        // STNkd(64)  :189-224
  L_E_C2, L_E_C3,                                              // encoder conv2/3 :263-264
  L_HEAD0, L_HEAD1, L_HEAD2, L_HEAD3, L_HEAD4,                 // cls: fc1,fc2,fc3 ; seg: conv1g,conv1p,conv2,conv3,conv4
  L_COUNT
};

struct cg_net {
  cg_ctx *ctx;
  int kind, n_out;
  float *blob_dev;
  size_t blob_floats;
  cg_layer L[L_COUNT];
  // engine-1 (tcgen05) bf16 hi/lo operand images of each trunk's 128->1024, 64->128 and (STNkd) 64->64
  // layers, built at create time: [W3 | W2 | W1]; indices 0 = STN3d, 1 = STNkd, 2 = encoder
  void *tc_img[3];
  int tc_f16_ok[3];   // 128->1024 weights fit fp16 (|w| < 65504): the 2-pass engine may be used for this trunk
};

// How the first kernel of each trunk obtains its (N,6) input rows.
struct cg_input_src {
  const float *x_direct;   // (B,N,6) float32, or nullptr
  const double *cloud_xyz; // (M,3)
  const double *cloud_nrm; // (M,3)
  const double *poses;     // (B,4,4)
  const int32_t *ids;      // (B,N)
  const double *mean;      // (6) or nullptr
  const double *stdv;      // (6) or nullptr
  int M;
};

struct cg_trunk_args {
  cg_input_src in;
  int B, N;
  const float *T3;   // (B,9) or nullptr
  cg_layer l0;       // 6 -> 64  (+ReLU)
  int stage1_mode;   // 0 none, 1 shared 64->64 (+ReLU), 2 per-candidate 64x64 matrix (no bias / ReLU)
  cg_layer l1;
  const float *T64;  // (B,64,64) when stage1_mode == 2
  cg_layer l2;       // 64 -> 128 (+ReLU)
  cg_layer l3;       // 128 -> 1024
  const void *tc_img; // tcgen05 operand images [W3 | W2 | W1 | W3 fp16] of l3 / l2 / l1 (engines 1, 2)
  int tc_f16_ok;      // engine 2 allowed for this trunk (else it runs the 3-pass kernel)
  int relu3;
  uint32_t *gmax_keys;  // (B,1024) order-preserving keys, zero-initialised by the launcher
  float *pf_out;        // (B,N,64) stage-1 output (PointNetSeg point feature) or nullptr
  unsigned long long *dbg;  // optional per-CTA cycle counters (CG_TRUNK_DEBUG=1), else nullptr
  uint32_t *ovf_flag;       // engine 3: set to 1 when an activation had to be clamped to the fp16 range (or nullptr)
  int exp_flags;            // timing experiments only (CG_TRUNK_EXP bitmask, results become wrong): 1 = no W3 copies,
                            // 2 = max warps skip the TMEM reads, 4 = front warps skip their math
};

int cg_trunk_launch_simt(cg_ctx *ctx, const cg_trunk_args &a);
int cg_trunk_launch_tc(cg_ctx *ctx, const cg_trunk_args &a);
int cg_trunk_launch_p(cg_ctx *ctx, const cg_trunk_args &a);   // persistent single-pass kernel (engine 3)
size_t cg_tc_image_bytes();
// Wt3 [128][1024], Wt2 [64][128], Wt1 [64][64] or nullptr (folded fp32, k-major rows, host)
int cg_tc_prepare(cg_ctx *ctx, const float *Wt3, const float *Wt2, const float *Wt1, void *dst_dev, int *f16_ok);

// Y[M][N] = act(X[M][K] @ Wt[K][N] + bias[(row / bias_row_div)][N])
// x_is_keys: X holds order-preserving uint keys (output of a trunk) to be decoded on load.
int cg_linear_launch(cg_ctx *ctx, const float *X, int M, int K, const float *Wt, const float *bias,
                     int N, int relu, int bias_row_div, int x_is_keys, float *Y);
// tensor-core FC path (cg_linear_tc.cu)
int cg_linear_tc_register(cg_ctx *ctx, const float *Wt_dev, const float *Wt_host, int K, int N);
void cg_linear_tc_unregister(const float *Wt_dev);
int cg_linear_tc_try(cg_ctx *ctx, const float *X, int M, int K, const float *Wt, const float *bias, int N, int relu,
                     int bias_row_div, int x_is_keys, float *Y);
int cg_softmax_launch(cg_ctx *ctx, const float *logits, int B, int C, float *probs, int32_t *label);
int cg_nunocs_post_launch(cg_ctx *ctx, const float *logits, int P, int bins, float *coords,
                          float *conf_z, int32_t *out_bins);
