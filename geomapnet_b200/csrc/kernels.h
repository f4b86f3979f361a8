// Internal launcher declarations for the MapNet B200 kernels.
#pragma once
#include "common.cuh"

namespace mapnet {

// Geometry of one convolution (NHWC activations, weights [Co][KH][KW][Ci] "KRSC").
struct ConvGeom {
  int B, Hi, Wi, Ci, Ho, Wo, Co, KH, KW, stride, pad;
  long long M_out() const { return (long long)B * Ho * Wo; }
  long long M_in() const { return (long long)B * Hi * Wi; }
  int Kdim() const { return KH * KW * Ci; }
};

// ---- bn.cu -------------------------------------------------------------------
template <typename T>
int launch_channel_sums(int mode, const T* a, const T* zmask, const T* y, const T* yd, long long M, int C,
                        float* partials, int* nblk_out, cudaStream_t st);
int launch_bn_fwd_finalize(const float* partials, int nblk, int C, long long M, const float* gamma,
                           const float* beta, float* run_mean, float* run_var, float* mean_out,
                           float* invstd_out, float* scale, float* shift, int training, cudaStream_t st);
template <typename T>
int launch_bn_apply(const T* y, const float* scale, const float* shift, int res_mode, const T* res,
                    const float* scale2, const float* shift2, T* z, long long M, int C, int relu, cudaStream_t st);
template <typename T>
int launch_stem_pool(const T* y, const float* scale, const float* shift, T* z, uint8_t* amax, int B, int H,
                     int W, int Ho, int Wo, int C, cudaStream_t st);
template <typename T>
int launch_stem_pool_bwd(const T* dz, const uint8_t* amax, const T* y, const float* scale, const float* shift,
                         T* g, int B, int H, int W, int Ho, int Wo, int C, cudaStream_t st);
int launch_bn_bwd_finalize(const float* partials, int nblk, int nacc, int which, int C, long long M,
                           const float* gamma, const float* mean, const float* invstd, float* dgamma,
                           float* dbeta, float* coef, cudaStream_t st);
template <typename T>
int launch_bn_bwd_apply(const T* dout, const T* zmask, const T* y, const float* coef, T* dy, const T* yd,
                        const float* coefd, T* dyd, T* gout, long long M, int C, cudaStream_t st);

// ---- conv_simt.cu: fp32 CUDA-core implicit GEMM (strict-parity path) -----------
template <typename T>
int launch_conv_simt_fprop(const ConvGeom& g, const T* x, const float* w_krsc, const T* residual, T* y, cudaStream_t st);
template <typename T>
int launch_conv_simt_dgrad(const ConvGeom& g, const T* dy, const float* w_dg /*[Ci][KH][KW][Co]*/, const T* residual, T* dx, cudaStream_t st);
template <typename T>
int launch_conv_simt_wgrad(const ConvGeom& g, const T* x, const T* dy, float* dw_krsc /*zeroed, accumulated*/, cudaStream_t st);

// ---- conv_tc.cu: tcgen05 / TMA implicit GEMM (bf16 tensor-core path) ------------
struct TcConvPlan;   // opaque: tensor maps + tap tables for one conv at one batch size
int tc_plan_create(TcConvPlan** out, const ConvGeom& g, int kind /*0 fprop,1 dgrad,2 wgrad*/, const bf16* act_in,
                   const bf16* wmat, const bf16* act_in2, void* out_ptr);
void tc_plan_destroy(TcConvPlan* p);
int tc_conv_run(TcConvPlan* p, const bf16* residual, cudaStream_t st);
int tc_selftest(int which, float* max_err_out, cudaStream_t st);

// ---- layout.cu -----------------------------------------------------------------
struct WeightDesc {   // one conv's weight in the flat parameter buffer and in the packed matrices
  long long p_off;    // float offset in params_flat ([Co,Ci_real,KH,KW] torch layout)
  long long k_off;    // element offset in the KRSC / dgrad packed buffers
  int Co, Ci, Ci_real, KH, KW;   // Ci = packed (padded) channels per tap
  int im2col_k;       // >0: stem conv stored as [Co][im2col_k] with k=(kh*KW+kw)*Ci_real+ci
};
template <typename TW>
int launch_pack_weights(const WeightDesc* d_descs, int nconv, const float* params, TW* w_krsc, TW* w_dg,
                        int max_elems, cudaStream_t st);
int launch_unpack_wgrads(const WeightDesc* d_descs, int nconv, const float* dw_krsc, float* grads,
                         int max_elems, cudaStream_t st);
template <typename T>
int launch_stem_im2col(const float* x_nchw, T* A, int B, int H, int W, int Ho, int Wo, int Kpad, cudaStream_t st);

// ---- head.cu -------------------------------------------------------------------
template <typename T>
int launch_gap(const T* z, float* feat, int B, int HW, int C, cudaStream_t st);
int launch_fc_fwd(const float* in, const float* w, const float* bias, float* out, int B, int In, int Out,
                  int relu, const float* mask_or_null, cudaStream_t st);
int launch_dropout_mask(float* mask, long long n, float p, unsigned long long seed, unsigned long long offset,
                        cudaStream_t st);
int launch_head_bwd(const float* dpred, const float* feat, const float* fcpre /*pre-relu [B,F]*/, const float* mask,
                    const float* w_fc, const float* w6, float* dh, float* dfeat, float* g_wfc, float* g_bfc,
                    float* g_w6, float* g_b6, int B, int C, int F, int filter_nans, cudaStream_t st);
template <typename T>
int launch_gap_bwd(const float* dfeat, T* dz, int B, int HW, int C, cudaStream_t st);

// ---- loss.cu -------------------------------------------------------------------
int launch_loss(int mode, const float* pred, const float* targ, int N, int Tp, int Tt, const float* s4,
                float* loss, float* dpred, float* ds4, cudaStream_t st);

// ---- adam.cu -------------------------------------------------------------------
int launch_sqnorm(const float* g, long long n, float* partials, float* out_sq, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                float eps, float wd, float bc1, float bc2, float gscale, const float* sqnorm_or_null,
                float max_norm, cudaStream_t st);

}  // namespace mapnet
