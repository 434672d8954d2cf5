// RegressionLoss (src/losses/regression.py:11-75) — the class the reference registers as `stereo_const` (virtual stereo
// consistency, src/core/handlers.py:151-198) and `depth_regr` (proxy depth regression, :201-259): SURVEY 8f rank 4.
//   p, t = invert ? to_inv(pred), to_inv(target) : pred, target          (to_inv: (x > 0) / max(x, eps), geometry.py:86-90)
//   diff = |p - t|;   e = diff | log(1 + diff) | berHu(diff; delta = 0.2 max(diff) over ALL elements, masked or not)
//   loss = sum(mask e) / sum(mask)                                        (mask = NULL: ones)
// Differentiable in pred AND target (stereo_const hands two network outputs), including berHu's dynamic threshold: the gradient
// through `diff.max()` goes to the maximal element(s), evenly (torch's full-reduction max). Everything is reduced through
// per-block partials combined by one block in a fixed order: deterministic, no atomics.
#include "stv_common.cuh"

namespace stv {

constexpr int RG_NT = 256, RG_MAXB = 1024;

struct RegrScalars {   // ws header (doubles): filled by the forward, read by the backward
    double max_diff, mask_sum, loss_sum, ddelta_sum, ties;
};

__device__ __forceinline__ float rg_inv(float x) { return x > 0.f ? 1.f/fmaxf(x, STV_EPS32) : 0.f; }
// d to_inv(x)/dx: (x > 0) * d(1/clamp(x, eps))/dx = -(1/x^2) where x >= eps (clamp passes the gradient at the bound), else 0
__device__ __forceinline__ float rg_dinv(float x) { return (x > 0.f && x >= STV_EPS32) ? -1.f/(x*x) : 0.f; }

__device__ __forceinline__ double rg_block_sum(double v, double* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = RG_NT/2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    const double r = sh[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double rg_block_max(double v, double* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = RG_NT/2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]); __syncthreads(); }
    const double r = sh[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ float rg_diff(const float* __restrict__ pred, const float* __restrict__ tgt, long long q, int invert) {
    const float p = __ldg(pred + q), t = __ldg(tgt + q);
    return fabsf(invert ? rg_inv(p) - rg_inv(t) : p - t);
}

// pass 1: per-block max(diff) and sum(mask) -> part[block] / part[MAXB + block]
__global__ void __launch_bounds__(RG_NT) regr_pass1_kernel(long long n, int invert, const float* __restrict__ pred, const float* __restrict__ tgt,
                                                           const float* __restrict__ mask, double* __restrict__ part) {
    __shared__ double sh[RG_NT];
    double mx = 0.0, ms = 0.0;
    for (long long q = (long long)blockIdx.x*RG_NT + threadIdx.x; q < n; q += (long long)gridDim.x*RG_NT) {
        mx = fmax(mx, (double)rg_diff(pred, tgt, q, invert));
        ms += mask ? (double)__ldg(mask + q) : 1.0;
    }
    mx = rg_block_max(mx, sh);
    ms = rg_block_sum(ms, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = mx; part[RG_MAXB + blockIdx.x] = ms; }
}

__global__ void __launch_bounds__(RG_NT) regr_reduce1_kernel(int nb, const double* __restrict__ part, RegrScalars* __restrict__ sc) {
    __shared__ double sh[RG_NT];
    double mx = 0.0, ms = 0.0;
    for (int q = threadIdx.x; q < nb; q += RG_NT) { mx = fmax(mx, part[q]); ms += part[RG_MAXB + q]; }
    mx = rg_block_max(mx, sh);
    ms = rg_block_sum(ms, sh);
    if (threadIdx.x == 0) { sc->max_diff = mx; sc->mask_sum = ms; }
}

// e(diff) and its partial derivatives; delta only matters for berHu
__device__ __forceinline__ void rg_error(int loss, float diff, float delta, float& e, float& de_ddiff, float& de_ddelta) {
    de_ddelta = 0.f;
    if (loss == STV_REGR_L1) { e = diff; de_ddiff = 1.f; }
    else if (loss == STV_REGR_LOG_L1) { e = log1pf(diff); de_ddiff = 1.f/(1.f + diff); }
    else if (diff <= delta) { e = diff; de_ddiff = 1.f; }
    else {
        const float den = 2.f*delta + STV_EPS32, num = diff*diff + delta*delta;
        e = num/den; de_ddiff = 2.f*diff/den; de_ddelta = (2.f*delta*den - 2.f*num)/(den*den);
    }
}

// pass 2: err map, per-block sum(mask e), sum(mask de/ddelta), number of maximal elements
__global__ void __launch_bounds__(RG_NT) regr_pass2_kernel(long long n, int loss, int invert, const float* __restrict__ pred,
                                                           const float* __restrict__ tgt, const float* __restrict__ mask,
                                                           const RegrScalars* __restrict__ sc, float* __restrict__ err, double* __restrict__ part) {
    __shared__ double sh[RG_NT];
    const float mx = (float)sc->max_diff, delta = 0.2f*mx;
    double ls = 0.0, ds = 0.0, ties = 0.0;
    for (long long q = (long long)blockIdx.x*RG_NT + threadIdx.x; q < n; q += (long long)gridDim.x*RG_NT) {
        const float diff = rg_diff(pred, tgt, q, invert), m = mask ? __ldg(mask + q) : 1.f;
        float e, d1, d2;
        rg_error(loss, diff, delta, e, d1, d2);
        if (err) err[q] = m*e;
        ls += (double)(m*e); ds += (double)(m*d2);
        if (diff == mx) ties += 1.0;
    }
    ls = rg_block_sum(ls, sh); ds = rg_block_sum(ds, sh); ties = rg_block_sum(ties, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = ls; part[RG_MAXB + blockIdx.x] = ds; part[2*RG_MAXB + blockIdx.x] = ties; }
}

__global__ void __launch_bounds__(RG_NT) regr_reduce2_kernel(int nb, const double* __restrict__ part, RegrScalars* __restrict__ sc,
                                                             float* __restrict__ loss_out) {
    __shared__ double sh[RG_NT];
    double ls = 0.0, ds = 0.0, ties = 0.0;
    for (int q = threadIdx.x; q < nb; q += RG_NT) { ls += part[q]; ds += part[RG_MAXB + q]; ties += part[2*RG_MAXB + q]; }
    ls = rg_block_sum(ls, sh); ds = rg_block_sum(ds, sh); ties = rg_block_sum(ties, sh);
    if (threadIdx.x == 0) { sc->loss_sum = ls; sc->ddelta_sum = ds; sc->ties = ties; *loss_out = (float)(ls/sc->mask_sum); }
}

__global__ void __launch_bounds__(RG_NT) regr_bwd_kernel(long long n, int loss, int invert, const float* __restrict__ pred,
                                                         const float* __restrict__ tgt, const float* __restrict__ mask,
                                                         const RegrScalars* __restrict__ sc, const float* __restrict__ grad_loss,
                                                         float* __restrict__ g_pred, float* __restrict__ g_tgt) {
    const long long q = (long long)blockIdx.x*RG_NT + threadIdx.x;
    if (q >= n) return;
    const float mx = (float)sc->max_diff, delta = 0.2f*mx;
    const float g = __ldg(grad_loss)/(float)sc->mask_sum;
    const float p = __ldg(pred + q), t = __ldg(tgt + q);
    const float pi = invert ? rg_inv(p) : p, ti = invert ? rg_inv(t) : t;
    const float d = pi - ti, diff = fabsf(d), m = mask ? __ldg(mask + q) : 1.f;
    float e, d1, d2;
    rg_error(loss, diff, delta, e, d1, d2);
    float gd = g*m*d1;                                                       // d loss / d diff_q through its own error
    if (loss == STV_REGR_BERHU && diff == mx) gd += g*0.2f*(float)(sc->ddelta_sum/sc->ties);   // ... and through the dynamic threshold
    const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    if (g_pred) g_pred[q] = gd*s*(invert ? rg_dinv(p) : 1.f);
    if (g_tgt) g_tgt[q] = -gd*s*(invert ? rg_dinv(t) : 1.f);
}

}  // namespace stv

using namespace stv;

extern "C" size_t stv_regr_workspace_bytes(void) { return (sizeof(RegrScalars) + 3*RG_MAXB*sizeof(double) + 255)/256*256; }

static int regr_check(long long n, int loss, const float* pred, const float* tgt, const void* ws, size_t ws_bytes, const char* what) {
    STV_REQUIRE(n > 0 && pred && tgt, "%s: empty input / NULL pointer", what);
    STV_REQUIRE(loss == STV_REGR_L1 || loss == STV_REGR_LOG_L1 || loss == STV_REGR_BERHU, "%s: bad loss %d (l1 | log_l1 | berhu)", what, loss);
    if (!ws || ws_bytes < stv_regr_workspace_bytes()) { set_error("%s: workspace too small", what); return STV_E_WORKSPACE; }
    return STV_OK;
}

extern "C" int stv_regr_fwd(long long n, int loss, int invert, const float* pred, const float* target, const float* mask, float* loss_out,
                            float* err, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = regr_check(n, loss, pred, target, ws, ws_bytes, "stv_regr_fwd")) return rc;
    STV_REQUIRE(loss_out, "stv_regr_fwd: NULL output");
    cudaStream_t st = (cudaStream_t)stream;
    RegrScalars* sc = (RegrScalars*)ws;
    double* part = (double*)((char*)ws + sizeof(RegrScalars));
    long long nbl = (n + RG_NT*4 - 1)/(RG_NT*4);
    const int nb = (int)(nbl < 1 ? 1 : (nbl > RG_MAXB ? RG_MAXB : nbl));
    regr_pass1_kernel<<<nb, RG_NT, 0, st>>>(n, invert, pred, target, mask, part);
    count_launch();
    if (int rc = check_launch("regr_pass1_kernel")) return rc;
    regr_reduce1_kernel<<<1, RG_NT, 0, st>>>(nb, part, sc);
    count_launch();
    if (int rc = check_launch("regr_reduce1_kernel")) return rc;
    regr_pass2_kernel<<<nb, RG_NT, 0, st>>>(n, loss, invert, pred, target, mask, sc, err, part);
    count_launch();
    if (int rc = check_launch("regr_pass2_kernel")) return rc;
    regr_reduce2_kernel<<<1, RG_NT, 0, st>>>(nb, part, sc, loss_out);
    count_launch();
    return check_launch("regr_reduce2_kernel");
}

extern "C" int stv_regr_bwd(long long n, int loss, int invert, const float* pred, const float* target, const float* mask,
                            const float* grad_loss, float* g_pred, float* g_target, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = regr_check(n, loss, pred, target, ws, ws_bytes, "stv_regr_bwd")) return rc;
    STV_REQUIRE(grad_loss && (g_pred || g_target), "stv_regr_bwd: NULL pointer");
    regr_bwd_kernel<<<(unsigned)((n + RG_NT - 1)/RG_NT), RG_NT, 0, (cudaStream_t)stream>>>(n, loss, invert, pred, target, mask,
                                                                                             (const RegrScalars*)ws, grad_loss, g_pred, g_target);
    count_launch();
    return check_launch("regr_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// The two pointwise regularisers (registry keys `disp_occ`, `disp_mask`): a mean over all elements, same partial-sum machinery.
//   STV_PWREG_MEAN     OccReg  (src/regularizers/occlusion.py:9-40):  loss = sign * mean(x)
//   STV_PWREG_BCE_ONE  MaskReg (src/regularizers/mask.py:11-30):      loss = F.binary_cross_entropy(x, 1) = mean(-max(log x, -100)),
//                      gradient (x - 1) / max((1 - x) x, 1e-12) / n as ATen computes it
// ---------------------------------------------------------------------------------------------------------------------
namespace stv {

__global__ void __launch_bounds__(RG_NT) pwreg_fwd_kernel(long long n, int kind, const float* __restrict__ x, double* __restrict__ part) {
    __shared__ double sh[RG_NT];
    double a = 0.0;
    for (long long q = (long long)blockIdx.x*RG_NT + threadIdx.x; q < n; q += (long long)gridDim.x*RG_NT) {
        const float v = __ldg(x + q);
        a += kind == STV_PWREG_MEAN ? (double)v : (double)(-fmaxf(logf(v), -100.f));
    }
    a = rg_block_sum(a, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

__global__ void __launch_bounds__(RG_NT) pwreg_reduce_kernel(int nb, const double* __restrict__ part, double scale, float* __restrict__ out) {
    __shared__ double sh[RG_NT];
    double a = 0.0;
    for (int q = threadIdx.x; q < nb; q += RG_NT) a += part[q];
    a = rg_block_sum(a, sh);
    if (threadIdx.x == 0) *out = (float)(a*scale);
}

__global__ void __launch_bounds__(RG_NT) pwreg_bwd_kernel(long long n, int kind, float sign, const float* __restrict__ x,
                                                          const float* __restrict__ grad_loss, float* __restrict__ g) {
    const long long q = (long long)blockIdx.x*RG_NT + threadIdx.x;
    if (q >= n) return;
    const float go = __ldg(grad_loss)/(float)n;
    if (kind == STV_PWREG_MEAN) { g[q] = sign*go; return; }
    const float v = __ldg(x + q);
    g[q] = go*(v - 1.f)/fmaxf((1.f - v)*v, 1e-12f);
}

}  // namespace stv

extern "C" int stv_pwreg_fwd(long long n, int kind, float sign, const float* x, float* loss, void* ws, size_t ws_bytes, void* stream) {
    STV_REQUIRE(n > 0 && x && loss, "stv_pwreg_fwd: empty input / NULL pointer");
    STV_REQUIRE(kind == STV_PWREG_MEAN || kind == STV_PWREG_BCE_ONE, "stv_pwreg_fwd: bad kind %d", kind);
    if (!ws || ws_bytes < stv_regr_workspace_bytes()) { set_error("stv_pwreg_fwd: workspace too small"); return STV_E_WORKSPACE; }
    double* part = (double*)((char*)ws + sizeof(RegrScalars));
    long long nbl = (n + RG_NT*4 - 1)/(RG_NT*4);
    const int nb = (int)(nbl < 1 ? 1 : (nbl > RG_MAXB ? RG_MAXB : nbl));
    pwreg_fwd_kernel<<<nb, RG_NT, 0, (cudaStream_t)stream>>>(n, kind, x, part);
    count_launch();
    if (int rc = check_launch("pwreg_fwd_kernel")) return rc;
    pwreg_reduce_kernel<<<1, RG_NT, 0, (cudaStream_t)stream>>>(nb, part, (kind == STV_PWREG_MEAN ? (double)sign : 1.0)/(double)n, loss);
    count_launch();
    return check_launch("pwreg_reduce_kernel");
}

extern "C" int stv_pwreg_bwd(long long n, int kind, float sign, const float* x, const float* grad_loss, float* g, void* stream) {
    STV_REQUIRE(n > 0 && x && grad_loss && g, "stv_pwreg_bwd: empty input / NULL pointer");
    STV_REQUIRE(kind == STV_PWREG_MEAN || kind == STV_PWREG_BCE_ONE, "stv_pwreg_bwd: bad kind %d", kind);
    pwreg_bwd_kernel<<<(unsigned)((n + RG_NT - 1)/RG_NT), RG_NT, 0, (cudaStream_t)stream>>>(n, kind, sign, x, grad_loss, g);
    count_launch();
    return check_launch("pwreg_bwd_kernel");
}
