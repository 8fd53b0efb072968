// Fused fit / predict drivers: the host side of the hot path, below the C ABI.
//
//  pb_laplace_fit   LaplaceGP.weight + precision  (probit/approximators.py:204-210,265-277;
//                   f_LA probit/implicit/Laplace.py:4-9; newton_solver + fwd_solver
//                   probit/implicit/solvers.py:7-25 with jaxopt's stopping rule)
//  pb_vb_fit        VBGP.weight + precision       (approximators.py:332-339; f_VB VB.py:4-16)
//  pb_predict*      Approximator.predict          (approximators.py:154-180)
//
// Newton step.  The reference forms J = diag(h) K - I densely (2N^3 flops through jax.jacobian)
// and LU-solves it ((2/3)N^3).  With W = -h >= 0, s = sqrt(W), b = W f + g, B = I + s s^T o K:
//     w+ = w - J^{-1}(g - w) = (I + W K)^{-1} b = b - s o B^{-1} (s o (K b))
// (Rasmussen & Williams Alg. 3.1), which needs one Cholesky of the SPD, well-conditioned B
// (N^3/3 flops) and never divides by W.  The iterates agree with the LU form to ~1e-14
// (tests/test_oracle_fit.py) and the iteration count is identical.
#include "likelihood.cuh"
#include "workspace.cuh"
#include <cstdlib>
#include <vector>

namespace pb {

namespace {

// two partial sums per block -> partial[2*block + k]
__device__ __forceinline__ void write_partials(double a, double b, double* partial) {
    a = block_sum<256>(a);
    b = block_sum<256>(b);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
    }
}

__global__ void __launch_bounds__(256)
finalize_kernel(const double* __restrict__ partial, int nblk, double* out0, double* out1) {
    double a = 0, b = 0;
    for (int i = threadIdx.x; i < nblk; i += 256) {
        a += partial[2 * i];
        b += partial[2 * i + 1];
    }
    a = block_sum<256>(a);
    b = block_sum<256>(b);
    if (threadIdx.x == 0) {
        if (out0) *out0 = a;
        if (out1) *out1 = b;
    }
}

// Newton prep (Laplace.py:4-9 + its derivative): s = sqrt(W), b = W f + g with W = -h.
// partials: sum ll, number of data with W < 0 or NaN.
__global__ void __launch_bounds__(256)
laplace_prep_kernel(lik::Params p, const double* __restrict__ cut, const double* __restrict__ f,
                    const void* __restrict__ y, int64_t n, double neg_floor, double* __restrict__ s,
                    double* __restrict__ b, double* __restrict__ Wout, double* __restrict__ partial) {
    PB_LIK_SMEM(sc);
    lik::stage_cutpoints(p, cut, sc);
    double sum_ll = 0, bad = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double fi = f[i];
        const lik::Out o = lik::eval(p, fi, y, i, sc);
        double W = -o.h;
        // log(Z + 1e-10) is not log-concave where Z <~ 1e-10 (a datum 5.5 .. 8.7 sigma outside its interval): h is
        // positive there, ~1e-9 in the far tail and up to ~9 / sigma^2 around Z ~ eps.  The reference's LU Newton step
        // simply carries such data along.  Here: a curvature in [neg_floor, 0) is treated as 0 (same fixed point, the
        // step solves a marginally different linearisation); anything below keeps its sign and sends the step down
        // the signed-Cholesky path (indefinite_newton_solve).  `bad` counts those data (NaN included).
        if (!(W >= neg_floor)) bad += 1.0;
        else W = W > 0.0 ? W : 0.0;
        s[i] = sqrt(W > 0.0 ? W : 0.0);
        b[i] = fma(W, fi, o.g);
        Wout[i] = W;
        sum_ll += o.ll;
    }
    write_partials(sum_ll, bad, partial);
}

// precision p = -h(f) (approximators.py:274-276); partials: sum ll(f), f.w
__global__ void __launch_bounds__(256)
posterior_stats_kernel(lik::Params p, const double* __restrict__ cut, const double* __restrict__ f,
                       const void* __restrict__ y, const double* __restrict__ w, int64_t n, double* __restrict__ prec,
                       double* __restrict__ partial) {
    PB_LIK_SMEM(sc);
    lik::stage_cutpoints(p, cut, sc);
    double sum_ll = 0, ftw = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double fi = f[i];
        const lik::Out o = lik::eval(p, fi, y, i, sc);
        if (prec) prec[i] = -o.h;
        sum_ll += o.ll;
        ftw = fma(fi, w[i], ftw);
    }
    write_partials(sum_ll, ftw, partial);
}

// VB right-hand side (VB.py:11-15): r = f + sigma * g(f)
__global__ void __launch_bounds__(256)
vb_rhs_kernel(lik::Params p, const double* __restrict__ cut, const double* __restrict__ f, const void* __restrict__ y,
              int64_t n, double* __restrict__ r) {
    PB_LIK_SMEM(sc);
    lik::stage_cutpoints(p, cut, sc);
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double fi = f[i];
        const lik::Out o = lik::eval(p, fi, y, i, sc);
        r[i] = fma(p.sigma, o.g, fi);
    }
}

__global__ void __launch_bounds__(256)
mul_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = a[i] * b[i];
}

__global__ void __launch_bounds__(256)
sqrt_kernel(const double* __restrict__ a, int64_t n, double neg_floor, double* __restrict__ out,
            double* __restrict__ partial) {
    double bad = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = a[i];
        if (!(v >= neg_floor)) bad += 1.0;
        out[i] = sqrt(v > 0.0 ? v : 0.0);      // a precision in [neg_floor, 0) is a datum that carries no information
    }
    write_partials(0.0, bad, partial);
}

__global__ void __launch_bounds__(256)
fill_kernel(double* __restrict__ a, int64_t n, double v) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) a[i] = v;
}

// w+ = b - s o c ; partial: ||w+ - w||^2   (solvers.py:24 + jaxopt's error)
__global__ void __launch_bounds__(256)
newton_update_kernel(const double* __restrict__ b, const double* __restrict__ s, const double* __restrict__ c,
                     const double* __restrict__ w, int64_t n, double* __restrict__ wn, double* __restrict__ partial) {
    double e2 = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = fma(-s[i], c[i], b[i]);
        const double d = v - w[i];
        e2 = fma(d, d, e2);
        wn[i] = v;
    }
    write_partials(e2, 0.0, partial);
}

__global__ void __launch_bounds__(256)
diff_norm_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ partial) {
    double e2 = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double d = a[i] - b[i];
        e2 = fma(d, d, e2);
    }
    write_partials(e2, 0.0, partial);
}

// var[i] = kss - sum_j V[i][j]^2 ; one warp per row
__global__ void __launch_bounds__(256)
row_sumsq_kernel(const double* __restrict__ V, int64_t rows, int64_t cols, int64_t ld, double kss,
                 double* __restrict__ var) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const double* p = V + r * ld;
    double s = 0;
    const int64_t c2 = cols >> 1;
    for (int64_t j = lane; j < c2; j += 32) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(p) + j);
        s = fma(v.x, v.x, s);
        s = fma(v.y, v.y, s);
    }
    if ((cols & 1) && lane == 0) s = fma(p[cols - 1], p[cols - 1], s);
    s = warp_sum(s);
    if (lane == 0) var[r] = kss - s;
}

// ---- preconditioned CG on B x = c, B = I + s s^T o K, preconditioner = a stale Cholesky factor ----
// (used by the later Newton iterations, where W changes little: ~15 iterations of one symv + two trsv
// replace an N^3/3 factorisation; the solution is driven to 1e-13 relative residual, so the iterates
// agree with the direct solve far inside the 1e-8 parity tolerance.)

// e_i = clamp(s_fac_i / s_i): M^{-1} = E B_fac^{-1} E is SPD for any positive diagonal E
__global__ void __launch_bounds__(256)
pcg_scale_kernel(const double* __restrict__ s, const double* __restrict__ sf, int64_t n, double* __restrict__ e) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double v = (s[i] > 0.0 && sf[i] > 0.0) ? sf[i] / s[i] : 1.0;
        e[i] = fmin(fmax(v, 0.25), 4.0);
    }
}

// r = c, y = 0; partial: ||c||^2
__global__ void __launch_bounds__(256)
pcg_init_kernel(const double* __restrict__ c, int64_t n, double* __restrict__ r, double* __restrict__ y,
                double* __restrict__ partial) {
    double a = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = c[i];
        r[i] = v;
        y[i] = 0.0;
        a = fma(v, v, a);
    }
    write_partials(a, 0.0, partial);
}

// r = c - (y + s o t) with t = K (s o y); partials: ||c||^2, ||r||^2
__global__ void __launch_bounds__(256)
pcg_warm_kernel(const double* __restrict__ c, const double* __restrict__ y, const double* __restrict__ s,
                const double* __restrict__ t, int64_t n, double* __restrict__ r, double* __restrict__ partial) {
    double a = 0, b = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double ci = c[i];
        const double ri = ci - fma(s[i], t[i], y[i]);
        r[i] = ri;
        a = fma(ci, ci, a);
        b = fma(ri, ri, b);
    }
    write_partials(a, b, partial);
}

// out = a o b; optional partial: sum r o out (the r.z product of PCG)
__global__ void __launch_bounds__(256)
pcg_mul_dot_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ r, int64_t n,
                   double* __restrict__ out, double* __restrict__ partial) {
    double d = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double v = a[i] * b[i];
        out[i] = v;
        if (r) d = fma(r[i], v, d);
    }
    write_partials(d, 0.0, partial);
}

// q = p + s o v (= B p); partial: p.q
__global__ void __launch_bounds__(256)
pcg_bp_kernel(const double* __restrict__ p, const double* __restrict__ s, const double* __restrict__ v, int64_t n,
              double* __restrict__ q, double* __restrict__ partial) {
    double d = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double qi = fma(s[i], v[i], p[i]);
        q[i] = qi;
        d = fma(p[i], qi, d);
    }
    write_partials(d, 0.0, partial);
}

// alpha = rz / pBp (device scalars); y += alpha p; r -= alpha q; partial: ||r||^2
__global__ void __launch_bounds__(256)
pcg_update_kernel(const double* __restrict__ sc_rz, const double* __restrict__ sc_pbp, const double* __restrict__ p,
                  const double* __restrict__ q, int64_t n, double* __restrict__ y, double* __restrict__ r,
                  double* __restrict__ partial) {
    const double alpha = *sc_rz / *sc_pbp;
    double d = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        y[i] = fma(alpha, p[i], y[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        d = fma(ri, ri, d);
    }
    write_partials(d, 0.0, partial);
}

// p = z + (rz_new / rz_old) p
__global__ void __launch_bounds__(256)
pcg_dir_kernel(const double* __restrict__ sc_new, const double* __restrict__ sc_old, const double* __restrict__ z,
               int64_t n, double* __restrict__ p) {
    const double beta = *sc_new / *sc_old;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) p[i] = fma(beta, p[i], z[i]);
}

// ---- evidence-gradient helpers (oracle/gradients.py) ----
// s2_i = 1/2 V_i d3_i with V_i = (1 - Binv_ii) / W_i = diag((K^-1 + W)^-1);
// partials: [0] Gaussian d(-Psi)/dsigma terms  sum(-1/sigma + (y-f)^2/sigma^3 + V_i/sigma^3), [1] #(W_i <= 0)
__global__ void __launch_bounds__(256)
grad_s2_kernel(const double* __restrict__ neg_binv_diag, const double* __restrict__ W, const double* __restrict__ d3,
               const double* __restrict__ f, const void* __restrict__ y, int gaussian, double sigma, int64_t n,
               double neg_floor, double kss, double* __restrict__ s2, double* __restrict__ Vout,
               double* __restrict__ partial) {
    double acc = 0, bad = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double w = W[i];
        if (!(w >= neg_floor)) bad += 1.0;
        // neg_binv_diag = -sum_k U_ik^2 = -(B^-1)_ii.  A datum whose curvature was clamped to 0 (Z << eps: every
        // derivative of log(Z + eps) is ~1e-9 there, so V only ever multiplies ~0) gets the prior variance, the
        // bound V_i <= K_ii.
        const double V = w > 0.0 ? (1.0 + neg_binv_diag[i]) / w : kss;
        Vout[i] = V;
        s2[i] = 0.5 * V * d3[i];
        if (gaussian) {
            const double r = reinterpret_cast<const double*>(y)[i] - f[i];
            acc += -1.0 / sigma + (r * r + V) / (sigma * sigma * sigma);
        }
    }
    write_partials(acc, bad, partial);
}

// Ordinal-probit likelihood parameters (oracle/gradients.py ordinal_parameter_partials): per datum
//   t = dll/dphi + 1/2 V dh/dphi + uvec dg/dphi   for phi = sigma, the lower and the upper cutpoint,
// accumulated per block in shared memory: slot 0 = sigma, slot 1 + j = cutpoint b_j.  partial[block][J + 2].
__global__ void __launch_bounds__(256)
ordinal_param_grad_kernel(const double* __restrict__ f, const long long* __restrict__ y, const double* __restrict__ cut,
                          int J, double sigma, double eps, const double* __restrict__ V, const double* __restrict__ uvec,
                          int64_t n, double* __restrict__ partial) {
    PB_LIK_SMEM(sc);
    __shared__ double acc[lik::MAX_CUT + 2];
    for (int i = threadIdx.x; i <= J; i += 256) sc[i] = cut[i];
    lik::stage_tables(sc);
    const double* tbl = sc + lik::TBL_OFF;
    for (int i = threadIdx.x; i < J + 2; i += 256) acc[i] = 0.0;
    __syncthreads();
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        long long yi = y[i];
        yi = yi < 0 ? 0 : (yi >= J ? J - 1 : yi);
        const double b1 = sc[yi], b2 = sc[yi + 1], fi = f[i];
        const bool fin1 = b1 != -INFINITY, fin2 = b2 != INFINITY;
        const double z1 = fin1 ? (b1 - fi) / sigma : 0.0, z2 = fin2 ? (b2 - fi) / sigma : 0.0;
        const double u = ((fin2 ? lik::norm_cdf(z2, tbl) : 1.0) - (fin1 ? lik::norm_cdf(z1, tbl) : 0.0)) + eps;
        const double A = (fin1 ? lik::norm_z_pdf(z1) : 0.0) / u, B = (fin2 ? lik::norm_z_pdf(z2) : 0.0) / u;
        const double L1 = -A, L2 = B;
        const double L11 = z1 * A - A * A, L12 = A * B, L22 = -z2 * B - B * B;
        const double L111 = A * (1.0 - z1 * z1) + 3.0 * z1 * A * A - 2.0 * A * A * A;
        const double L112 = -z1 * A * B + 2.0 * A * A * B;
        const double L122 = -2.0 * A * B * B - z2 * A * B;
        const double L222 = B * (z2 * z2 - 1.0) + 3.0 * z2 * B * B + 2.0 * B * B * B;
        const double S = L11 + 2.0 * L12 + L22;
        const double s1 = 1.0 / sigma, s2i = s1 * s1, s3i = s2i * s1;
        const double hv = 0.5 * V[i], uv = uvec[i];
        const double t_sigma = -(z1 * L1 + z2 * L2) * s1
                               + hv * (-(2.0 * S + z1 * (L111 + 2.0 * L112 + L122) + z2 * (L112 + 2.0 * L122 + L222)) * s3i)
                               + uv * (((L1 + L2) + z1 * (L11 + L12) + z2 * (L12 + L22)) * s2i);
        const double t_lower = L1 * s1 + hv * ((L111 + 2.0 * L112 + L122) * s3i) + uv * (-(L11 + L12) * s2i);
        const double t_upper = L2 * s1 + hv * ((L112 + 2.0 * L122 + L222) * s3i) + uv * (-(L12 + L22) * s2i);
        atomicAdd(&acc[0], t_sigma);
        if (fin1) atomicAdd(&acc[1 + yi], t_lower);
        if (fin2) atomicAdd(&acc[2 + yi], t_upper);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < J + 2; i += 256) partial[(int64_t)blockIdx.x * (J + 2) + i] = acc[i];
}

__global__ void __launch_bounds__(256)
sum_columns_kernel(const double* __restrict__ partial, int nblk, int width, double* __restrict__ out) {
    for (int c = threadIdx.x; c < width; c += 256) {
        double a = 0.0;
        for (int b = 0; b < nblk; ++b) a += partial[(int64_t)b * width + c];
        out[c] = a;
    }
}

// out = a - b
__global__ void __launch_bounds__(256)
sub_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = a[i] - b[i];
}

// partials: a.b and c.d
__global__ void __launch_bounds__(256)
dot2_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
            const double* __restrict__ d, int64_t n, double* __restrict__ partial) {
    double p = 0, q = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        p = fma(a[i], b[i], p);
        q = fma(c[i], d[i], q);
    }
    write_partials(p, q, partial);
}

int check_problem(const pb_problem* prob) {
    PB_CHECK(prob != nullptr, PB_ERR_INVALID, "null problem");
    PB_CHECK(prob->n >= 1 && prob->D >= 1, PB_ERR_INVALID, "problem needs n >= 1 and D >= 1");
    PB_CHECK(prob->X != nullptr && prob->y != nullptr, PB_ERR_INVALID, "problem data missing");
    return PB_OK;
}

int bind(const pb_problem* prob, void* workspace, int64_t workspace_bytes, Ws& ws) {
    PB_TRY(check_problem(prob));
    ws.L = make_layout(prob->n, prob->D);
    PB_CHECK(workspace != nullptr && workspace_bytes >= ws.L.total, PB_ERR_INVALID,
             "workspace too small: need %lld bytes, got %lld", (long long)ws.L.total, (long long)workspace_bytes);
    PB_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, PB_ERR_INVALID, "workspace must be 256-byte aligned");
    ws.base = reinterpret_cast<uint8_t*>(workspace);
    return PB_OK;
}

int build_gram(cudaStream_t st, const pb_problem* prob, const Ws& ws) {
    const int Df = feature_dim(prob->kernel, prob->D);
    PB_TRY(features(st, prob->kernel, prob->X, prob->n, prob->D, prob->D, ws.Z(), prob->n));
    if (ws.dist)      // this rank's rows of K only, straight from the features: K[lo:hi, :] = k(Z[lo:hi], Z)
        return gram_cross(st, prob->kernel, ws.Z() + ws.dist->lo, ws.dist->hi - ws.dist->lo, ws.Z(), prob->n, Df, prob->n,
                          prob->n, ws.K(), ws.L.ld, nullptr);
    return gram_sym(st, prob->kernel, ws.Z(), prob->n, Df, prob->n, ws.K(), ws.L.ld, nullptr, 0.0);
}

int finalize(cudaStream_t st, const Ws& ws, unsigned nblk, double* out0, double* out1) {
    finalize_kernel<<<1, 256, 0, st>>>(ws.partial(), (int)nblk, out0, out1); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int read_scalars(cudaStream_t st, const Ws& ws, double* host, int32_t* info_host) {
    PB_CUDA(cudaMemcpyAsync(host, ws.scalars(), S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (info_host) PB_CUDA(cudaMemcpyAsync(info_host, ws.info(), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

int K_times(cudaStream_t st, const Ws& ws, int64_t n, const double* x, double* y, double* symv_scratch);

// posterior mean f = K w, precision, sum ll, f.w  -> device scalars
int posterior_stats(cudaStream_t st, const pb_problem* prob, const lik::Params& lp, const Ws& ws, const double* w,
                    double* prec_out) {
    const int64_t n = prob->n;
    PB_TRY(K_times(st, ws, n, w, ws.vec(V_F), nullptr));
    const unsigned nb = vec_blocks(n);
    posterior_stats_kernel<<<nb, 256, 0, st>>>(lp, prob->lik.cutpoints, ws.vec(V_F), prob->y, w, n, prec_out,
                                               ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return finalize(st, ws, nb, ws.scalars() + S_SUMLL, ws.scalars() + S_FTW);
}

// y = K x for the Newton / CG iterations: the half-traffic symv when its scratch is available, else the row-wise gemv.
int K_times(cudaStream_t st, const Ws& ws, int64_t n, const double* x, double* y, double* symv_scratch) {
    if (ws.dist) {
        // rows of K sharded over the ranks: each streams its n/G rows (the product is HBM bound, so 1/G of the time),
        // then ONE in-place all-gather of 8 n bytes enqueued on the same stream puts the same y on every rank
        const DistCtx& d = *ws.dist;
        PB_TRY(gemv(st, ws.K(), d.hi - d.lo, n, ws.L.ld, x, y + d.lo));
        return comm_allgather(d.comm, st, y, d.nloc_max);
    }
    if (symv_scratch) return symv_lower(st, ws.K(), n, ws.L.ld, x, y, symv_scratch);
    return gemv(st, ws.K(), n, n, ws.L.ld, x, y);
}

// Factor a I + s s^T o (K + jitter I) into ws.B() (lower) and fill the solve workspace.
int factor_matrix(cudaStream_t st, const Ws& ws, int64_t n, const double* s, double a, double jitter) {
    PB_TRY(sym_transform(st, ws.K(), n, ws.L.ld, s, a, jitter, ws.B(), ws.L.ld));
    return potrf(st, ws.B(), n, ws.L.ld, ws.potrf_ws(), pb_potrf_workspace_bytes(n), ws.info());
}

// B = I + s s^T o (K + jitter I), factor in place, logdet -> device scalar
int factor_B(cudaStream_t st, const Ws& ws, int64_t n, const double* s, double jitter) {
    PB_TRY(factor_matrix(st, ws, n, s, 1.0, jitter));
    return logdet_chol(st, ws.B(), n, ws.L.ld, ws.scalars() + S_LOGDET);
}

// Solve B(s) y = c by PCG; `precondition(rz_slot)` must put z = M^{-1} r into V_Z (r in V_R) and r.z into the
// device scalar rz_slot.  c is left intact; the solution lands in V_Y.  With `warm` the iteration starts from the
// vector already in V_Y (one extra symv for the initial residual).  *iters = iterations used, or -1 if the
// residual did not reach its target within `maxit` (the caller then factors B).  Synchronises `st` once per
// iteration (readback of the residual norm).
template <class Precond>
int pcg_run(cudaStream_t st, const Ws& ws, int64_t n, const double* s, const double* c, int maxit, double tol_abs,
            bool warm, int* iters, Precond&& precondition, double* symv_scratch = nullptr) {
    // stop at ||r|| <= max(tol_abs, 1e-15 ||c||): the absolute target comes from the Newton tolerance (cg_target), the
    // relative floor is what FP64 can deliver
    auto converged = [&](const double* host) {
        return host[S_RR] <= std::max(tol_abs * tol_abs, 1e-30 * host[S_R0]);
    };
    const unsigned nb = vec_blocks(n);
    auto Kmul = [&](const double* in, double* out) -> int { return K_times(st, ws, n, in, out, symv_scratch); };
    double* sc = ws.scalars();
    double host[S_COUNT];
    *iters = -1;
    if (warm) {
        // r = c - B y0: u = s o y0, t = K u, r = c - (y0 + s o t); partials: ||c||^2, ||r||^2
        pcg_mul_dot_kernel<<<nb, 256, 0, st>>>(s, ws.vec(V_Y), nullptr, n, ws.vec(V_U), ws.partial()); pb::note_launch();
        PB_TRY(Kmul(ws.vec(V_U), ws.vec(V_T)));
        pcg_warm_kernel<<<nb, 256, 0, st>>>(c, ws.vec(V_Y), s, ws.vec(V_T), n, ws.vec(V_R), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, sc + S_R0, sc + S_RR));
        PB_TRY(read_scalars(st, ws, host, nullptr));
        if (!(host[S_RR] == host[S_RR])) return PB_OK;
        if (converged(host)) { *iters = 0; return PB_OK; }
    } else {
        pcg_init_kernel<<<nb, 256, 0, st>>>(c, n, ws.vec(V_R), ws.vec(V_Y), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, sc + S_R0, nullptr));
    }
    int cur = 0;
    PB_TRY(precondition(sc + S_RZ0));
    PB_CUDA(cudaMemcpyAsync(ws.vec(V_P), ws.vec(V_Z), n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (int j = 1; j <= maxit; ++j) {
        pcg_mul_dot_kernel<<<nb, 256, 0, st>>>(s, ws.vec(V_P), nullptr, n, ws.vec(V_U), ws.partial()); pb::note_launch();
        PB_TRY(Kmul(ws.vec(V_U), ws.vec(V_T)));
        pcg_bp_kernel<<<nb, 256, 0, st>>>(ws.vec(V_P), s, ws.vec(V_T), n, ws.vec(V_Q), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, sc + S_PBP, nullptr));
        pcg_update_kernel<<<nb, 256, 0, st>>>(sc + (cur ? S_RZ1 : S_RZ0), sc + S_PBP, ws.vec(V_P), ws.vec(V_Q), n,
                                              ws.vec(V_Y), ws.vec(V_R), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, sc + S_RR, nullptr));
        PB_TRY(read_scalars(st, ws, host, nullptr));
        if (!(host[S_RR] == host[S_RR]) || !(host[S_PBP] > 0.0)) return PB_OK;       // breakdown: let the caller factor
        if (converged(host)) { *iters = j; return PB_OK; }
        PB_TRY(precondition(sc + (cur ? S_RZ0 : S_RZ1)));
        pcg_dir_kernel<<<nb, 256, 0, st>>>(sc + (cur ? S_RZ0 : S_RZ1), sc + (cur ? S_RZ1 : S_RZ0), ws.vec(V_Z), n,
                                           ws.vec(V_P)); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    return PB_OK;
}

// PCG preconditioned with the Cholesky factor currently in ws.B() (built for the s stored in V_SF):
// M^{-1} = E B_fac^{-1} E with E = clamp(s_fac / s), SPD for any positive diagonal E.
int pcg_solve(cudaStream_t st, const Ws& ws, int64_t n, const double* s, const double* c, int maxit, double tol,
              int* iters) {     // tol: absolute residual target (cg_target)
    const unsigned nb = vec_blocks(n);
    const int64_t ld = ws.L.ld;
    pcg_scale_kernel<<<nb, 256, 0, st>>>(s, ws.vec(V_SF), n, ws.vec(V_E)); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    auto precondition = [&](double* rz_slot) -> int {      // z = E B_fac^{-1} E r ; rz = r.z
        pcg_mul_dot_kernel<<<nb, 256, 0, st>>>(ws.vec(V_E), ws.vec(V_R), nullptr, n, ws.vec(V_U), ws.partial()); pb::note_launch();
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_U), ws.vec(V_Z)));
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_Z), ws.vec(V_U)));
        pcg_mul_dot_kernel<<<nb, 256, 0, st>>>(ws.vec(V_E), ws.vec(V_U), ws.vec(V_R), n, ws.vec(V_Z), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        return finalize(st, ws, nb, rz_slot, nullptr);
    };
    return pcg_run(st, ws, n, s, c, maxit, tol, false, iters, precondition);
}

// ---- Nystrom preconditioner: Newton steps without any N^3 factorisation ----
// With r landmarks I = {0, q, 2q, ...}, q = floor(n / r) (strided, so that inputs sorted along some coordinate are
// still covered evenly),
//     K ~ K_NI (K_II + delta I)^{-1} K_IN,      M = I + S K_NI (K_II + delta I)^{-1} K_IN S   (S = diag(s))
//     M^{-1} = I - S K_NI A^{-1} K_IN S,        A = K_II + delta I + K_IN S^2 K_NI            (Woodbury)
// G = K_IN S (r x N, the landmark rows of the Gram matrix already in HBM, scaled) is kept, so one application is
// z = r - G^T A^{-1} G r: two r x N matvecs and two r x r triangular solves (< 1 ms at N = 65536, r = 4096) beside
// the 3 ms symv of the CG step itself.  A is rebuilt for every Newton step (one r x r x N lower GEMM on the tensor cores, 33 ms, and an
// r x r Cholesky).  Measured on the north-star workload: ~25 CG iterations per Newton step to 1e-13, against 147+
// unpreconditioned; EQ and long lengthscales need fewer (the kernel is closer to low rank), short lengthscales
// need few without any help.  ws.B() is idle until the final factorisation, so everything lives there.
struct Nystrom {
    int64_t r = 0, lda = 0, pws_bytes = 0;
    int64_t ldg = 0, ncols = 0, col0 = 0;   // G covers columns [col0, col0 + ncols) (all of them on one GPU, the rank's shard otherwise)
    double* Zl = nullptr;     // features of the landmarks, feature-major (multi-GPU only)
    double* G = nullptr;      // r x ldg : K_IN S
    double* A = nullptr;      // r x lda
    double* pws = nullptr;    // potrf workspace of A
    double* t = nullptr;      // r
    double* u = nullptr;      // r
    double* symv = nullptr;   // scratch of the half-traffic symv (null when ld is not padded to 64)
    double* gt = nullptr;     // gemv_t partials: splits x ld
    int64_t stride = 1;       // landmark i is training point i * stride
    void* oz = nullptr;       // slicing scratch of the INT8 G G^T (r rows x oz_kb columns at a time); null: FP64 DMMA
    int64_t oz_kb = 0, oz_bytes = 0;
};
constexpr int64_t NYSTROM_OZ_KB = 16384;

int64_t nystrom_rank(int64_t n) {
    long long r = opt_nystrom_rank();
    if (r < 0) r = std::min<long long>(4096, std::max<long long>(256, n / 16));
    r = std::min<long long>(std::min<long long>(r, n / 4), 32768);                 // grid.y of nystrom_prep_kernel
    return r >= 256 ? r / 256 * 256 : r / 64 * 64;                                 // small n (tests, sharded toy sizes): 64-granular
}

// Layout inside the (idle) factor region.  Returns the doubles used; `base` may be null (size query).
int64_t nystrom_carve(Nystrom& ny, double* base, int64_t n, int64_t ld, int64_t ldg, int Dfmax, bool sharded) {
    int64_t used = 0;
    auto take = [&](int64_t doubles) { double* o = base ? base + used : nullptr; used += round_up(doubles, 32); return o; };
    ny.lda = round_up(ny.r, 16);
    ny.pws_bytes = pb_potrf_workspace_bytes(ny.r);
    ny.ldg = ldg;
    ny.G = take(ny.r * ldg);
    ny.A = take(ny.r * ny.lda);
    ny.pws = take(ny.pws_bytes / 8 + 1);
    ny.t = take(ny.r);
    ny.u = take(ny.r);
    ny.gt = take(gemv_t_splits(ny.r) * ldg);
    ny.stride = n / ny.r;
    if (sharded) ny.Zl = take((int64_t)Dfmax * ny.r);
    else if (ld >= round_up(n, 64)) ny.symv = take(symv_lower_scratch_doubles(n));
    if (ozaki_enabled(n) && ny.r >= 1024 && (ldg & 1) == 0) {      // G G^T (r x r x n, rebuilt every Newton step) on the INT8 tensor cores
        ny.oz_kb = std::min<int64_t>(NYSTROM_OZ_KB, ldg / 64 * 64);
        ny.oz_bytes = ny.oz_kb >= 1024 ? ozaki_scratch_bytes(ny.r, ny.oz_kb) : 0;
        ny.oz = ny.oz_bytes ? reinterpret_cast<void*>(take(ny.oz_bytes / 8 + 1)) : nullptr;
        if (!base) ny.oz = nullptr;
    }
    return used;
}

Nystrom nystrom_layout(const Ws& ws, int64_t n) {
    Nystrom ny;
    ny.r = nystrom_rank(n);
    if (ny.r <= 0) return ny;
    const bool sharded = ws.dist != nullptr;
    ny.ncols = sharded ? ws.dist->hi - ws.dist->lo : n;
    ny.col0 = sharded ? ws.dist->lo : 0;
    const int64_t used = nystrom_carve(ny, ws.B(), n, ws.L.ld, sharded ? ws.dist->nloc_max : ws.L.ld, ws.L.Dfmax, sharded);
    if (used > ws.L.B_doubles) ny.r = 0;      // does not fit (tiny n): disabled
    return ny;
}

// G[i][j] = K[l_i][j] s[j] for the landmark rows l_i = i * stride; A[i][j] = K[l_i][l_j] + delta [i == j]
__global__ void __launch_bounds__(256)
nystrom_prep_kernel(const double* __restrict__ K, int64_t ldk, const double* __restrict__ s, int64_t n, int64_t r,
                    int64_t stride, double delta, double* __restrict__ G, double* __restrict__ A, int64_t lda) {
    const int64_t i = blockIdx.y;
    const double* row = K + i * stride * ldk;
    for (int64_t j = blockIdx.x * 256ll + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256) {
        G[i * ldk + j] = row[j] * s[j];
        if (j < r) A[i * lda + j] = row[j * stride] + (i == j ? delta : 0.0);
    }
}

// z = r - sum_k vpart[k][.] ; partial: r.z
__global__ void __launch_bounds__(256)
nystrom_z_kernel(const double* __restrict__ r, const double* __restrict__ vpart, int splits, int64_t ldp, int64_t n,
                 double* __restrict__ z, double* __restrict__ partial) {
    double d = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double ri = r[i];
        double v = 0.0;
        for (int k = 0; k < splits; ++k) v += vpart[k * ldp + i];
        const double zi = ri - v;
        z[i] = zi;
        d = fma(ri, zi, d);
    }
    write_partials(d, 0.0, partial);
}

// Zl[d][i] = Z[d][i * stride]: the landmarks' features, feature-major with leading dimension r
__global__ void __launch_bounds__(256)
gather_landmarks_kernel(const double* __restrict__ Z, int64_t ldz, int Df, int64_t r, int64_t stride,
                        double* __restrict__ Zl) {
    for (int64_t e = blockIdx.x * 256ll + threadIdx.x; e < (int64_t)Df * r; e += (int64_t)gridDim.x * 256) {
        const int64_t d = e / r, i = e % r;
        Zl[e] = Z[d * ldz + i * stride];
    }
}

// Builds A for the current s and factors it; *ok = false if A is not numerically SPD.
// Multi-GPU: G is split by columns — rank q holds G[:, lo_q:hi_q] = k(landmarks, X[lo_q:hi_q]) S, generated from the
// features, forms its r x r x (n/G) share of G G^T on the tensor cores, and ONE all-reduce of the r x r block
// (134 MB at r = 4096) completes A = K_II + delta I + G G^T on every rank; the r x r Cholesky is replicated.
// A (lower) += G G^T, G r x K: K-blocks of ny.oz_kb on the INT8 tensor cores (the preconditioner needs no more than a few
// digits, but the sliced product is exact to 2^-49 anyway), the tail and small problems on the DMMA kernel.
int nystrom_syrk(cudaStream_t st, const Nystrom& ny, int64_t K, const double* G, int64_t ldg) {
    int64_t k = 0;
    if (ny.oz && ny.oz_bytes > 0)
        for (; K - k >= 1024; ) {
            const int64_t kb = std::min<int64_t>(ny.oz_kb, (K - k) / 64 * 64);
            PB_TRY(ozaki_syrk_lower(st, ny.r, kb, 1.0, G + k, ldg, ny.A, ny.lda, ny.oz, ny.oz_bytes));
            k += kb;
        }
    if (k < K) PB_TRY(gemm_nt(st, ny.r, ny.r, K - k, 1.0, G + k, ldg, G + k, ldg, 1.0, ny.A, ny.lda, true));
    return PB_OK;
}

int nystrom_build(cudaStream_t st, const Ws& ws, const Nystrom& ny, int64_t n, const double* s, double delta, bool* ok,
                  const pb_problem* prob) {
    const int64_t ld = ws.L.ld;
    int32_t* info = ws.info() + 1;
    PB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    if (ws.dist) {
        const DistCtx& d = *ws.dist;
        const int Df = feature_dim(prob->kernel, prob->D);
        gather_landmarks_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(Df * ny.r, 256), 1024), 256, 0, st>>>(
            ws.Z(), n, Df, ny.r, ny.stride, ny.Zl); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        if (d.comm->rank == 0) PB_TRY(gram_sym(st, prob->kernel, ny.Zl, ny.r, Df, ny.r, ny.A, ny.lda, nullptr, delta));
        else PB_CUDA(cudaMemsetAsync(ny.A, 0, ny.r * ny.lda * sizeof(double), st));
        if (ny.ncols > 0) {
            PB_TRY(gram_cross(st, prob->kernel, ny.Zl, ny.r, ws.Z() + d.lo, ny.ncols, Df, ny.r, n, ny.G, ny.ldg, s + d.lo));
            PB_TRY(nystrom_syrk(st, ny, ny.ncols, ny.G, ny.ldg));
        }
        PB_TRY(comm_allreduce_sum(d.comm, st, ny.A, ny.r * ny.lda));
    } else {
        dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(n, 256), 64), (unsigned)ny.r);
        nystrom_prep_kernel<<<grid, 256, 0, st>>>(ws.K(), ld, s, n, ny.r, ny.stride, delta, ny.G, ny.A, ny.lda); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(nystrom_syrk(st, ny, n, ny.G, ld));
    }
    PB_TRY(potrf(st, ny.A, ny.r, ny.lda, ny.pws, ny.pws_bytes, info));
    int32_t info_host = 0;
    PB_CUDA(cudaMemcpyAsync(&info_host, info, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    *ok = info_host == 0;
    return PB_OK;
}

int nystrom_pcg(cudaStream_t st, const Ws& ws, const Nystrom& ny, int64_t n, const double* s, const double* c,
                int maxit, double tol, bool warm, int* iters) {
    const unsigned nb = vec_blocks(n);
    const int64_t ld = ws.L.ld;
    const int splits = gemv_t_splits(ny.r);
    auto precondition = [&](double* rz_slot) -> int {      // z = r - G^T A^{-1} G r ; rz = r.z
        if (ws.dist) {
            // column-sharded G: t = sum over ranks of G_q r_q (all-reduce of r doubles), the r x r solves are
            // replicated, z_q = r_q - G_q^T t is local, one all-gather returns z; r.z is then taken on the full
            // (replicated) vectors so every rank computes bit-identical CG scalars
            const DistCtx& d = *ws.dist;
            const unsigned nbl = vec_blocks(ny.ncols);
            if (ny.ncols > 0) PB_TRY(gemv(st, ny.G, ny.r, ny.ncols, ny.ldg, ws.vec(V_R) + d.lo, ny.t));
            else PB_CUDA(cudaMemsetAsync(ny.t, 0, ny.r * sizeof(double), st));
            PB_TRY(comm_allreduce_sum(d.comm, st, ny.t, ny.r));
            PB_TRY(trsv(st, ny.A, ny.r, ny.lda, ny.pws, false, ny.t, ny.u));
            PB_TRY(trsv(st, ny.A, ny.r, ny.lda, ny.pws, true, ny.u, ny.t));
            if (ny.ncols > 0) {
                PB_TRY(gemv_t_partial(st, ny.G, ny.r, ny.ncols, ny.ldg, ny.t, ny.gt, ny.ldg));
                nystrom_z_kernel<<<nbl, 256, 0, st>>>(ws.vec(V_R) + d.lo, ny.gt, splits, ny.ldg, ny.ncols, ws.vec(V_Z) + d.lo,
                                                      ws.partial()); pb::note_launch();
            }
            PB_TRY(comm_allgather(d.comm, st, ws.vec(V_Z), d.nloc_max));
            dot2_kernel<<<nb, 256, 0, st>>>(ws.vec(V_R), ws.vec(V_Z), ws.vec(V_R), ws.vec(V_Z), n, ws.partial()); pb::note_launch();
            PB_CUDA(cudaGetLastError());
            return finalize(st, ws, nb, rz_slot, nullptr);
        }
        PB_TRY(gemv(st, ny.G, ny.r, n, ld, ws.vec(V_R), ny.t));
        PB_TRY(trsv(st, ny.A, ny.r, ny.lda, ny.pws, false, ny.t, ny.u));
        PB_TRY(trsv(st, ny.A, ny.r, ny.lda, ny.pws, true, ny.u, ny.t));
        PB_TRY(gemv_t_partial(st, ny.G, ny.r, n, ld, ny.t, ny.gt, ld));
        nystrom_z_kernel<<<nb, 256, 0, st>>>(ws.vec(V_R), ny.gt, splits, ld, n, ws.vec(V_Z), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        return finalize(st, ws, nb, rz_slot, nullptr);
    };
    return pcg_run(st, ws, n, s, c, maxit, tol, warm, iters, precondition, ny.symv);
}

// Smallest curvature W = -h accepted as "zero" (pb_options.negative_curvature_tol, in units of 1/sigma^2).
double curvature_floor(const pb_problem* prob) {
    return -opts().negative_curvature_tol / (prob->lik.sigma * prob->lik.sigma);
}

// ---- Newton step with indefinite curvature: block elimination ----
// With W = S D S, S = |W|^1/2, D = diag(+-1):   (I + W K)^-1 = I - S (D + S K S)^-1 S K   (for D = I this is the B form).
// M = D + S K S is symmetric indefinite.  Order the data with non-negative curvature first (p of them, stable) and
// those with negative curvature last (m):
//     M = [ B+   E^T ]      B+ = I + S+ K++ S+  (always SPD: Cholesky, right-TRSM and GEMM on the tensor cores),
//         [ E    C0  ]      C0 = -I + S- K-- S-
//     Yt = E L+^-T,   C = C0 - Yt Yt^T  (m x m Schur complement, symmetric, in general INDEFINITE),
//     z1 = L+^-1 c1,  C x2 = c2 - Yt z1  (Gaussian elimination with partial pivoting),  x1 = L+^-T (z1 - Yt^T x2).
// This is the solution the reference's LU solve (solvers.py:24) returns whenever I + W K is non-singular, whatever
// the inertia of the Hessian — the reference walks through non-convex regions of log(Z + 1e-10) and still converges.
// The permuted matrix is generated from permuted features; p is made even with a decoupled dummy row so that the
// second block column stays 16-byte aligned for TMA.  Cost: one p^3/3 factorisation + m^3/3 elimination traffic for
// that Newton step — the rare path (small sigma or far-out cutpoints); m is the number of offending data.
__global__ void __launch_bounds__(1024)
partition_kernel(const double* __restrict__ W, int64_t n, double neg_floor, long long* __restrict__ perm,
                 long long* __restrict__ counts) {
    __shared__ long long cnt[1024];
    const int t = threadIdx.x;
    const int64_t chunk = (n + 1023) / 1024;
    const int64_t lo = min(n, (int64_t)t * chunk), hi = min(n, lo + chunk);
    long long c = 0;
    for (int64_t i = lo; i < hi; ++i) c += (W[i] >= neg_floor) ? 1 : 0;
    cnt[t] = c;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const long long v = t >= off ? cnt[t - off] : 0;
        __syncthreads();
        cnt[t] += v;
        __syncthreads();
    }
    const long long p = cnt[1023], pe = p + (p & 1);
    long long pos = cnt[t] - c, neg = lo - pos;
    for (int64_t i = lo; i < hi; ++i) {
        if (W[i] >= neg_floor) perm[pos++] = i;
        else perm[pe + neg++] = i;
    }
    if (t == 0) {
        if (p & 1) perm[p] = -1;         // dummy: s = 0, unit diagonal, decoupled
        counts[0] = p;
        counts[1] = pe;
    }
}

// permuted copies: features, sp = |W|^1/2, cp = sp o t   (index -1 = the dummy row)
__global__ void __launch_bounds__(256)
gather_perm_kernel(const long long* __restrict__ perm, int64_t np, const double* __restrict__ W, const double* __restrict__ t,
                   const double* __restrict__ Z, int64_t ldz, int Df, double* __restrict__ Zp, int64_t ldzp,
                   double* __restrict__ sp, double* __restrict__ cp) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < np; i += (int64_t)gridDim.x * 256) {
        const long long idx = perm[i];
        const double sv = idx >= 0 ? sqrt(fabs(W[idx])) : 0.0;
        sp[i] = sv;
        cp[i] = idx >= 0 ? sv * t[idx] : 0.0;
        for (int d = 0; d < Df; ++d) Zp[d * ldzp + i] = idx >= 0 ? Z[d * ldz + idx] : 0.0;
    }
}

__global__ void __launch_bounds__(256)
scatter_perm_kernel(const long long* __restrict__ perm, int64_t np, const double* __restrict__ xp, double* __restrict__ x) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < np; i += (int64_t)gridDim.x * 256) {
        const long long idx = perm[i];
        if (idx >= 0) x[idx] = xp[i];
    }
}

__global__ void __launch_bounds__(256)
abs_sqrt_kernel(const double* __restrict__ a, int64_t n, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) out[i] = sqrt(fabs(a[i]));
}

// A[i][i] += delta for i0 <= i < i1
__global__ void __launch_bounds__(256)
diag_add_kernel(double* __restrict__ A, int64_t ld, int64_t i0, int64_t i1, double delta) {
    for (int64_t i = i0 + blockIdx.x * 256ll + threadIdx.x; i < i1; i += (int64_t)gridDim.x * 256) A[i * ld + i] += delta;
}

// Gaussian elimination with partial pivoting on [C | rhs] (C m x m row-major, full storage), one column per pair of
// launches: (1) pivot search in column k + row swap, (2) elimination of the rows below.  Singular pivot -> *info.
__global__ void __launch_bounds__(1024)
lu_pivot_kernel(double* __restrict__ C, int64_t ld, int m, int k, double* __restrict__ rhs, int32_t* __restrict__ info,
                int32_t info_value) {
    __shared__ double bestv[32];
    __shared__ int besti[32];
    __shared__ int piv;
    double v = -1.0;
    int idx = k;
    for (int i = k + threadIdx.x; i < m; i += 1024) {
        const double a = fabs(C[(int64_t)i * ld + k]);
        if (a > v) { v = a; idx = i; }                          // NaN never wins: a NaN column ends with pivot 0 or NaN below
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bestv[threadIdx.x >> 5] = v; besti[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        v = bestv[threadIdx.x];
        idx = besti[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (threadIdx.x == 0) {
            piv = idx;
            if (!(v > 0.0)) atomicCAS(info, 0, info_value);
        }
    }
    __syncthreads();
    const int p = piv;
    if (p != k) {
        for (int j = k + threadIdx.x; j < m; j += 1024) {
            const double a = C[(int64_t)k * ld + j], b = C[(int64_t)p * ld + j];
            C[(int64_t)k * ld + j] = b;
            C[(int64_t)p * ld + j] = a;
        }
        if (threadIdx.x == 0) { const double a = rhs[k]; rhs[k] = rhs[p]; rhs[p] = a; }
    }
}

__global__ void __launch_bounds__(256)
lu_eliminate_kernel(double* __restrict__ C, int64_t ld, int m, int k, double* __restrict__ rhs) {
    const int i = k + 1 + blockIdx.x;
    const double f = C[(int64_t)i * ld + k] / C[(int64_t)k * ld + k];
    const double* pr = C + (int64_t)k * ld;
    double* row = C + (int64_t)i * ld;
    for (int j = k + 1 + threadIdx.x; j < m; j += 256) row[j] = fma(-f, pr[j], row[j]);
    if (threadIdx.x == 0) rhs[i] = fma(-f, rhs[k], rhs[i]);
}

// back substitution with the upper factor left in C: rhs <- U^-1 rhs (one CTA)
__global__ void __launch_bounds__(1024)
lu_backsub_kernel(const double* __restrict__ C, int64_t ld, int m, double* __restrict__ rhs) {
    __shared__ double xk;
    for (int k = m - 1; k >= 0; --k) {
        if (threadIdx.x == 0) {
            xk = rhs[k] / C[(int64_t)k * ld + k];
            rhs[k] = xk;
        }
        __syncthreads();
        const double x = xk;
        for (int i = threadIdx.x; i < k; i += 1024) rhs[i] = fma(-C[(int64_t)i * ld + k], x, rhs[i]);
        __syncthreads();
    }
}

// x (V_C, original order) = (D + S K S)^-1 (S o t), t = K b in V_T, signed W in V_G; V_S <- |W|^1/2.
// Overwrites the factor region, the potrf workspace and the PCG vector slots.
int indefinite_newton_solve(cudaStream_t st, const pb_problem* prob, const Ws& ws) {
    const int64_t n = prob->n;
    const int Df = feature_dim(prob->kernel, prob->D);
    const unsigned nb = vec_blocks(n + 1);
    long long* perm = reinterpret_cast<long long*>(ws.vec(V_U));
    long long* counts = reinterpret_cast<long long*>(ws.info() + 8);
    partition_kernel<<<1, 1024, 0, st>>>(ws.vec(V_G), n, curvature_floor(prob), perm, counts); pb::note_launch();
    long long host_counts[2] = {0, 0};
    PB_CUDA(cudaMemcpyAsync(host_counts, counts, sizeof(host_counts), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    const int64_t p = host_counts[0], pe = host_counts[1], np = n + (pe - p), m = np - pe;
    const int64_t ldm = round_up(np, 16);
    PB_CHECK(np * ldm <= ws.L.B_doubles, PB_ERR_INVALID, "indefinite Newton step: workspace too small");
    double *M = ws.B(), *Zp = ws.Zp(), *sp = ws.vec(V_SF);
    double *rhs = ws.vec(V_R), *z = ws.vec(V_Z), *tmp = ws.vec(V_E), *tx = ws.vec(V_X);
    gather_perm_kernel<<<nb, 256, 0, st>>>(perm, np, ws.vec(V_G), ws.vec(V_T), ws.Z(), n, Df, Zp, np, sp, rhs); pb::note_launch();
    abs_sqrt_kernel<<<nb, 256, 0, st>>>(ws.vec(V_G), n, ws.vec(V_S)); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    // M = I + sp sp^T o K[perm, perm] from the permuted features, then -2 on the diagonal of the negative block
    PB_TRY(gram_block(st, prob->kernel, Zp, np, Df, 0, np, 0, np, sp, 1.0, 0.0, M, ldm));
    if (m > 0) {
        diag_add_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(m, 256), 1024), 256, 0, st>>>(M, ldm, pe, np, -2.0); pb::note_launch();
        PB_CUDA(cudaGetLastError());
    }
    const int64_t pws_bytes = pb_potrf_workspace_bytes(n + 1);
    double* Yt = M + pe * ldm;            // m x pe
    double* Cb = Yt + pe;                 // m x m (full storage: the Schur complement is eliminated with row pivoting)
    if (pe > 0) PB_TRY(potrf(st, M, pe, ldm, ws.potrf_ws(), pws_bytes, ws.info()));
    else PB_CUDA(cudaMemsetAsync(ws.info(), 0, sizeof(int32_t), st));
    if (m > 0 && pe > 0) {
        PB_TRY(trsm_right_lt(st, M, pe, ldm, ws.potrf_ws(), Yt, m, ldm));
        PB_TRY(gemm_nt(st, m, m, pe, -1.0, Yt, ldm, Yt, ldm, 1.0, Cb, ldm, false));              // C = C0 - Yt Yt^T
    }
    const double* dinv1 = reinterpret_cast<const double*>(ws.potrf_ws());
    if (pe > 0) PB_TRY(trsv(st, M, pe, ldm, dinv1, false, rhs, z));                              // z1 = L+^-1 c1
    double* zhead = z;                   // z1 (later z1 - Yt^T x2) lives here
    if (m > 0) {
        const unsigned nbm = vec_blocks(m);
        double* c2 = tx + pe;            // right-hand side of the Schur system, overwritten by x2
        if (pe > 0) {
            PB_TRY(gemv(st, Yt, m, pe, ldm, z, tmp + pe));
            sub_kernel<<<nbm, 256, 0, st>>>(rhs + pe, tmp + pe, m, c2); pb::note_launch();
        } else {
            PB_CUDA(cudaMemcpyAsync(c2, rhs + pe, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        for (int64_t k = 0; k < m; ++k) {
            lu_pivot_kernel<<<1, 1024, 0, st>>>(Cb, ldm, (int)m, (int)k, c2, ws.info(), (int32_t)(pe + k + 1)); pb::note_launch();
            if (k + 1 < m) { lu_eliminate_kernel<<<(unsigned)(m - k - 1), 256, 0, st>>>(Cb, ldm, (int)m, (int)k, c2); pb::note_launch(); }
        }
        lu_backsub_kernel<<<1, 1024, 0, st>>>(Cb, ldm, (int)m, c2); pb::note_launch();          // x2
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaMemcpyAsync(rhs + pe, c2, m * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (pe > 0) {
            // z1 <- z1 - Yt^T x2, 1536 rows of Yt at a time (three partial slots), ping-pong z <-> tmp
            double* part = ws.vec(V_P);                   // V_P, V_Q, V_Y are consecutive slots
            double* cur = z;
            double* nxt = tmp;
            const unsigned nbp = vec_blocks(pe);
            for (int64_t i0 = 0; i0 < m; i0 += 3 * 512) {
                const int64_t rows = std::min<int64_t>(3 * 512, m - i0);
                PB_TRY(gemv_t_partial(st, Yt + i0 * ldm, rows, pe, ldm, rhs + pe + i0, part, ws.L.vec_stride));
                nystrom_z_kernel<<<nbp, 256, 0, st>>>(cur, part, gemv_t_splits(rows), ws.L.vec_stride, pe, nxt, ws.partial()); pb::note_launch();
                std::swap(cur, nxt);
            }
            zhead = cur;
        }
    }
    if (pe > 0) {
        double* out = zhead == z ? tmp : z;               // trsv: x must not alias rhs
        PB_TRY(trsv(st, M, pe, ldm, dinv1, true, zhead, out));                                   // x1 = L+^-T (z1 - Yt^T x2)
        PB_CUDA(cudaMemcpyAsync(rhs, out, pe * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    scatter_perm_kernel<<<nb, 256, 0, st>>>(perm, np, rhs, ws.vec(V_C)); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

bool pcg_enabled(int64_t n) { return n >= opt_pcg_min_n(); }

// Residual target of a CG Newton solve.  The step is w+ = b - s o x with B x = c, so a residual r leaves
// ||dw+|| <= max(s) ||B^-1|| ||r|| <= ||r|| / sigma  (B >= I; W = -h <= 1/sigma^2 for the probit family).
// Asking for ||dw+|| <= eta * tolerance (eta = "laplace_cg_tol", default 1e-2) keeps jaxopt's stopping test
// ||w+ - w|| <= tolerance faithful to 1 % and, measured at N = 32768, the returned weights within 2.5e-11 relative
// of the factor-every-step iterates (eta = 1e-1: 1.6e-10; eta <= 1e-3: 1.7e-11, the FP64 floor).
double cg_target(const pb_problem* prob, double tolerance) { return opt_cg_tol() * tolerance * prob->lik.sigma; }


}  // namespace

// ---- small helpers shared with dist.cu ----
// s = sqrt(max(p, 0)) into V_S; the count of materially negative / NaN precisions lands in the S_BAD scalar
int precision_sqrt(cudaStream_t st, const Ws& ws, const pb_problem* prob, const double* precision) {
    const unsigned nb = vec_blocks(prob->n);
    sqrt_kernel<<<nb, 256, 0, st>>>(precision, prob->n, curvature_floor(prob), ws.vec(V_S), ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return finalize(st, ws, nb, nullptr, ws.scalars() + S_BAD);
}

// var[i] = kss - ||V[i, :]||^2
int row_sumsq(cudaStream_t st, const double* V, int64_t rows, int64_t cols, int64_t ld, double kss, double* var) {
    if (rows <= 0) return PB_OK;
    row_sumsq_kernel<<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, st>>>(V, rows, cols, ld, kss, var); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

int64_t dist_nystrom_doubles(int64_t n, int64_t nloc_max, int Dfmax) {
    Nystrom ny;
    ny.r = nystrom_rank(n);
    if (ny.r <= 0) return 0;
    return nystrom_carve(ny, nullptr, n, round_up(n > 0 ? n : 1, 16), nloc_max, Dfmax, true);
}

}  // namespace pb

using namespace pb;

extern "C" int64_t pb_fit_workspace_bytes(int64_t n, int D) { return make_layout(n, D).total; }

extern "C" int pb_build_gram(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes) {
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    return build_gram(reinterpret_cast<cudaStream_t>(stream), prob, ws);
}

extern "C" int pb_build_features(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes) {
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    return features(reinterpret_cast<cudaStream_t>(stream), prob->kernel, prob->X, prob->n, prob->D, prob->D, ws.Z(),
                    prob->n);
}

extern "C" int pb_workspace_gram(void* workspace, int64_t n, int D, double** K, int64_t* ldk) {
    PB_CHECK(workspace && K && ldk, PB_ERR_INVALID, "workspace_gram: null argument");
    const Layout L = make_layout(n, D);
    *K = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + L.K);
    *ldk = L.ld;
    return PB_OK;
}

extern "C" int pb_laplace_fit(pb_stream_t stream, const pb_problem* prob, double tolerance, int32_t maxiter,
                              double jitter, int32_t final_factor, void* workspace, int64_t workspace_bytes,
                              double* weight, double* precision, double* posterior_mean,
                              pb_fit_result* result_host, const pb_options* options) {
    OptScope opt_scope(options);
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    return laplace_fit_impl(reinterpret_cast<cudaStream_t>(stream), prob, tolerance, maxiter, jitter, final_factor, ws,
                            weight, precision, posterior_mean, result_host);
}

int pb::laplace_fit_impl(cudaStream_t st, const pb_problem* prob, double tolerance, int32_t maxiter, double jitter,
                         int32_t final_factor, Ws& ws, double* weight, double* precision, double* posterior_mean,
                         pb_fit_result* result_host) {
    const bool sharded = ws.dist != nullptr;
    PB_CHECK(weight && precision && result_host, PB_ERR_INVALID, "laplace_fit: null output");
    PB_CHECK(!(sharded && final_factor), PB_ERR_INVALID, "laplace_fit: the sharded fit leaves the factorisation to pb_dist_predict");
    lik::Params lp;
    PB_TRY(lik::make_params(prob->lik, lp));
    const int64_t n = prob->n, ld = ws.L.ld;
    const unsigned nb = vec_blocks(n);
    *result_host = pb_fit_result{};

    PB_TRY(build_gram(st, prob, ws));
    PB_CUDA(cudaMemsetAsync(ws.scalars(), 0, S_COUNT * sizeof(double), st));
    PB_CUDA(cudaMemsetAsync(ws.info(), 0, 256, st));        // potrf is the only other writer: a CG-only fit must not read a stale word
    PB_CUDA(cudaMemsetAsync(ws.vec(V_W), 0, n * sizeof(double), st));       // z_init = zeros (approximators.py:268)

    double host[S_COUNT];
    int32_t info_host = 0;
    double error = INFINITY;
    int it = 0;
    double* w = ws.vec(V_W);
    double* wn = ws.vec(V_WN);
    bool have_factor = false;
    // Newton steps by Nystrom-preconditioned CG (no factorisation) until it fails once, then the factor path
    Nystrom ny;
    if (sharded || (pcg_enabled(n) && prob->lik.kind != PB_LIK_GAUSSIAN)) ny = nystrom_layout(ws, n);
    bool nystrom_live = ny.r > 0, nystrom_warm = false;
    // The sharded fit has no replicated factor to fall back on: its Newton steps are Nystrom-CG only.
    PB_CHECK(!sharded || nystrom_live, PB_ERR_UNSUPPORTED,
             "laplace_fit (multi-GPU): the Nystrom preconditioner is disabled or does not fit (n = %lld too small?)", (long long)n);
    while (error > tolerance && it < maxiter) {                             // jaxopt loop (solvers.py:13-14)
        if (it == 0) PB_CUDA(cudaMemsetAsync(ws.vec(V_F), 0, n * sizeof(double), st));   // K @ 0
        else PB_TRY(K_times(st, ws, n, w, ws.vec(V_F), nystrom_live && !have_factor ? ny.symv : nullptr));
        laplace_prep_kernel<<<nb, 256, 0, st>>>(lp, prob->lik.cutpoints, ws.vec(V_F), prob->y, n, curvature_floor(prob),
                                                ws.vec(V_S), ws.vec(V_B), ws.vec(V_G), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, ws.scalars() + S_SUMLL, ws.scalars() + S_BAD));
        PB_TRY(K_times(st, ws, n, ws.vec(V_B), ws.vec(V_T), nystrom_live && !have_factor ? ny.symv : nullptr));   // K b
        mul_kernel<<<nb, 256, 0, st>>>(ws.vec(V_S), ws.vec(V_T), n, ws.vec(V_C)); pb::note_launch();
        // materially negative curvature somewhere?  (8-byte readback; the count decides the solver below)
        PB_TRY(read_scalars(st, ws, host, nullptr));
        const bool indefinite = host[S_BAD] > 0;
        if (indefinite && sharded) {
            set_error("laplace_fit (multi-GPU): %d data have negative likelihood curvature (< %.3g) at iteration %d; the "
                      "signed-Cholesky Newton step is single-GPU only", (int)host[S_BAD], curvature_floor(prob), it + 1);
            return PB_ERR_NUMERIC;
        }
        // x = B^{-1} (s o K b).  Large n: CG with the Nystrom preconditioner, no factorisation at all.  If that
        // ever stalls, or below "laplace_pcg_min_n": the first iteration factors B; later ones reuse the last
        // factor as a PCG preconditioner (rescaled by s_fac/s) and refactor only if PCG stalls.
        const double* xsol = ws.vec(V_C);
        bool solved = false;
        if (indefinite) {
            PB_TRY(indefinite_newton_solve(st, prob, ws));      // x in V_C, V_S <- |W|^1/2
            have_factor = false;                                 // the factor region now holds the block-eliminated matrix
            nystrom_warm = false;
            result_host->factorizations += 1;
            solved = true;
        } else if (have_factor && prob->lik.kind == PB_LIK_GAUSSIAN) {
            // W = 1/sigma^2 does not depend on f: B is the matrix already factored, reuse it as is
            PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_C), ws.vec(V_X)));
            PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_X), ws.vec(V_C)));
            solved = true;
        } else if (have_factor && it >= 1 && pcg_enabled(n)) {
            int used = -1;
            PB_TRY(pcg_solve(st, ws, n, ws.vec(V_S), ws.vec(V_C), 48, cg_target(prob, tolerance), &used));
            if (used >= 0) {
                solved = true;
                xsol = ws.vec(V_Y);
                result_host->pcg_iterations += used;
            }
        } else if (nystrom_live && !have_factor) {
            int used = -1;
            bool ok = false;
            // delta regularises a numerically rank-deficient landmark block (EQ kernels: cond(K_II) > 1e16).  M is
            // SPD for any delta >= 0; delta perturbs the approximated K by ~(n/r) delta, i.e. the preconditioned
            // spectrum by (n/r) delta W ~ 0.04, while keeping cond(A) <~ 1e8.
            PB_TRY(nystrom_build(st, ws, ny, n, ws.vec(V_S), 1e-3 * prob->kernel.scale, &ok, prob));
            if (ok) PB_TRY(nystrom_pcg(st, ws, ny, n, ws.vec(V_S), ws.vec(V_C), 150, cg_target(prob, tolerance), nystrom_warm, &used));
            if (used >= 0) {
                solved = true;
                xsol = ws.vec(V_Y);
                nystrom_warm = true;
                result_host->pcg_iterations += used;
            } else {
                nystrom_live = false;
            }
        }
        PB_CHECK(solved || !sharded, PB_ERR_NUMERIC,
                 "laplace_fit (multi-GPU): Nystrom-preconditioned CG did not reach its target at Newton iteration %d", it + 1);
        if (!solved) {
            PB_TRY(factor_B(st, ws, n, ws.vec(V_S), 0.0));
            PB_CUDA(cudaMemcpyAsync(ws.vec(V_SF), ws.vec(V_S), n * sizeof(double), cudaMemcpyDeviceToDevice, st));
            have_factor = true;
            result_host->factorizations += 1;
            PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_C), ws.vec(V_X)));
            PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_X), ws.vec(V_C)));
        }
        newton_update_kernel<<<nb, 256, 0, st>>>(ws.vec(V_B), ws.vec(V_S), xsol, w, n, wn, ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, ws.scalars() + S_ERR2, nullptr));
        PB_TRY(read_scalars(st, ws, host, &info_host));
        ++it;
        result_host->iterations = it;
        result_host->info = info_host;
        if (info_host != 0) {
            if (indefinite)
                set_error("laplace_fit: %d data have negative likelihood curvature at iteration %d and the Newton matrix I + W K is "
                          "singular (or NaN) there: elimination broke down at column %d of the permuted matrix",
                          (int)host[S_BAD], it, info_host);
            else
                set_error("laplace_fit: Cholesky of I + W^1/2 K W^1/2 failed at column %d (iteration %d)", info_host, it);
            return PB_ERR_NUMERIC;
        }
        error = sqrt(host[S_ERR2]);
        result_host->error = error;
        if (!(error == error)) {
            set_error("laplace_fit: NaN iterate at iteration %d", it);
            return PB_ERR_NUMERIC;
        }
        double* tmp = w; w = wn; wn = tmp;
    }
    // LaplaceGP.precision (approximators.py:271-277) at the returned weight
    PB_TRY(posterior_stats(st, prob, lp, ws, w, precision));
    PB_CUDA(cudaMemcpyAsync(weight, w, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (posterior_mean)
        PB_CUDA(cudaMemcpyAsync(posterior_mean, ws.vec(V_F), n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (final_factor) {
        // chol(K + diag(1/p) + jitter I) of objective_LA (Laplace.py:24) in its B form:
        // sum log diag L_cov + 0.5 sum log p == sum log diag chol(I + s s^T o (K + jitter I))
        const unsigned nb2 = vec_blocks(n);
        sqrt_kernel<<<nb2, 256, 0, st>>>(precision, n, curvature_floor(prob), ws.vec(V_S), ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb2, nullptr, ws.scalars() + S_BAD));
        PB_TRY(factor_B(st, ws, n, ws.vec(V_S), jitter));
        result_host->factorizations += 1;
    }
    PB_TRY(read_scalars(st, ws, host, &info_host));
    result_host->sum_ll = host[S_SUMLL];
    result_host->ftw = host[S_FTW];
    result_host->logdet = final_factor ? host[S_LOGDET] : NAN;
    result_host->info = info_host;
    if (final_factor && (host[S_BAD] > 0 || info_host != 0)) {
        set_error("laplace_fit: final factorisation failed (bad precisions %d, potrf info %d)", (int)host[S_BAD],
                  info_host);
        return PB_ERR_NUMERIC;
    }
    return PB_OK;
}

extern "C" int pb_vb_fit(pb_stream_t stream, const pb_problem* prob, double tolerance, int32_t maxiter,
                         void* workspace, int64_t workspace_bytes, double* weight, double* precision,
                         double* posterior_mean, pb_fit_result* result_host, const pb_options* options) {
    OptScope opt_scope(options);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    PB_CHECK(weight && precision && result_host, PB_ERR_INVALID, "vb_fit: null output");
    lik::Params lp;
    PB_TRY(lik::make_params(prob->lik, lp));
    const int64_t n = prob->n, ld = ws.L.ld;
    const unsigned nb = vec_blocks(n);
    *result_host = pb_fit_result{};
    const double sigma = prob->lik.sigma;

    PB_TRY(build_gram(st, prob, ws));
    PB_CUDA(cudaMemsetAsync(ws.scalars(), 0, S_COUNT * sizeof(double), st));
    PB_CUDA(cudaMemsetAsync(ws.info(), 0, 256, st));
    // L = chol(sigma^2 I + K) (VB.py:10) — loop invariant, factored once (no jitter: raw-array path)
    PB_TRY(factor_matrix(st, ws, n, nullptr, sigma * sigma, 0.0));
    PB_TRY(logdet_chol(st, ws.B(), n, ld, ws.scalars() + S_LOGDET));
    PB_CUDA(cudaMemsetAsync(ws.vec(V_W), 0, n * sizeof(double), st));
    result_host->factorizations = 1;

    double host[S_COUNT];
    int32_t info_host = 0;
    PB_TRY(read_scalars(st, ws, host, &info_host));
    result_host->info = info_host;
    if (info_host != 0) {
        set_error("vb_fit: Cholesky of sigma^2 I + K failed at column %d", info_host);
        return PB_ERR_NUMERIC;
    }
    double error = INFINITY;
    int it = 0;
    double* w = ws.vec(V_W);
    double* wn = ws.vec(V_WN);
    while (error > tolerance && it < maxiter) {
        if (it == 0) PB_CUDA(cudaMemsetAsync(ws.vec(V_F), 0, n * sizeof(double), st));
        else PB_TRY(gemv(st, ws.K(), n, n, ld, w, ws.vec(V_F)));
        vb_rhs_kernel<<<nb, 256, 0, st>>>(lp, prob->lik.cutpoints, ws.vec(V_F), prob->y, n, ws.vec(V_B)); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_B), ws.vec(V_X)));     // cholesky_solve (VB.py:11)
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_X), wn));
        diff_norm_kernel<<<nb, 256, 0, st>>>(wn, w, n, ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, ws.scalars() + S_ERR2, nullptr));
        PB_TRY(read_scalars(st, ws, host, nullptr));
        ++it;
        error = sqrt(host[S_ERR2]);
        result_host->iterations = it;
        result_host->error = error;
        if (!(error == error)) {
            set_error("vb_fit: NaN iterate at iteration %d", it);
            return PB_ERR_NUMERIC;
        }
        double* tmp = w; w = wn; wn = tmp;
    }
    PB_TRY(posterior_stats(st, prob, lp, ws, w, nullptr));
    fill_kernel<<<nb, 256, 0, st>>>(precision, n, 1.0 / (sigma * sigma)); pb::note_launch();     // approximators.py:339
    PB_CUDA(cudaMemcpyAsync(weight, w, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (posterior_mean)
        PB_CUDA(cudaMemcpyAsync(posterior_mean, ws.vec(V_F), n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    PB_TRY(read_scalars(st, ws, host, nullptr));
    result_host->sum_ll = host[S_SUMLL];
    result_host->ftw = host[S_FTW];
    result_host->logdet = host[S_LOGDET];
    return PB_OK;
}

extern "C" int pb_predict_prepare(pb_stream_t stream, const pb_problem* prob, const double* precision,
                                  int32_t reuse_gram, void* workspace, int64_t workspace_bytes,
                                  int32_t* info_host, const pb_options* options) {
    OptScope opt_scope(options);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    PB_CHECK(precision != nullptr && info_host != nullptr, PB_ERR_INVALID, "predict_prepare: null argument");
    const int64_t n = prob->n;
    if (!reuse_gram) PB_TRY(build_gram(st, prob, ws));
    PB_CUDA(cudaMemsetAsync(ws.scalars(), 0, S_COUNT * sizeof(double), st));
    PB_CUDA(cudaMemsetAsync(ws.info(), 0, 256, st));
    const unsigned nb = vec_blocks(n);
    sqrt_kernel<<<nb, 256, 0, st>>>(precision, n, curvature_floor(prob), ws.vec(V_S), ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, nullptr, ws.scalars() + S_BAD));
    PB_TRY(factor_B(st, ws, n, ws.vec(V_S), 0.0));       // K + diag(1/p) (approximators.py:175) in B form
    double host[S_COUNT];
    PB_TRY(read_scalars(st, ws, host, info_host));
    if (host[S_BAD] > 0) {
        set_error("predict_prepare: %d precisions are negative or NaN", (int)host[S_BAD]);
        return PB_ERR_NUMERIC;
    }
    if (*info_host != 0) {
        set_error("predict_prepare: Cholesky failed at column %d", *info_host);
        return PB_ERR_NUMERIC;
    }
    return PB_OK;
}

extern "C" int64_t pb_predict_scratch_bytes(int64_t n, int D, int64_t chunk) {
    const int64_t ld = round_up(n > 0 ? n : 1, 16);
    return round_up(chunk * ld * 8, 256) + round_up(chunk * 2 * (int64_t)D * 8, 256) + round_up(64 * chunk * 8, 256) + 256;
}

extern "C" int pb_predict(pb_stream_t stream, const pb_problem* prob, const void* workspace, const double* weight,
                          const double* X_test, int64_t n_test, int64_t chunk, void* scratch,
                          int64_t scratch_bytes, double* mean, double* variance) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    PB_TRY(check_problem(prob));
    PB_CHECK(workspace != nullptr && weight != nullptr && mean != nullptr, PB_ERR_INVALID, "predict: null argument");
    PB_CHECK(n_test >= 0 && chunk >= 1, PB_ERR_INVALID, "predict: bad n_test/chunk");
    PB_CHECK(scratch != nullptr && scratch_bytes >= pb_predict_scratch_bytes(prob->n, prob->D, chunk), PB_ERR_INVALID,
             "predict: scratch too small");
    PB_CHECK((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, PB_ERR_INVALID, "predict: scratch must be 256-byte aligned");
    Ws ws;
    ws.L = make_layout(prob->n, prob->D);
    ws.base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(workspace));
    const int64_t n = prob->n, ld = ws.L.ld;
    const int D = prob->D, Df = feature_dim(prob->kernel, D);
    double* V = reinterpret_cast<double*>(scratch);
    double* Zs = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(scratch) + round_up(chunk * ld * 8, 256));
    double* mean_partial = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(Zs) + round_up(chunk * 2 * (int64_t)D * 8, 256));
    const double kss = prob->kernel.scale;    // kernel.elwise(x*, x*) of a stationary kernel (approximators.py:172)
    for (int64_t r0 = 0; r0 < n_test; r0 += chunk) {
        const int64_t m = n_test - r0 < chunk ? n_test - r0 : chunk;
        PB_TRY(features(st, prob->kernel, X_test + r0 * D, m, D, D, Zs, chunk));
        // mean = K_*f w (approximators.py:173,179): cross-covariance tiles generated in registers, never stored
        PB_TRY(gram_matvec(st, prob->kernel, Zs, m, ws.Z(), n, Df, chunk, n, weight, mean_partial, mean + r0));
        if (variance) {
            // var = k** - || L_B^{-1} (s o k_*) ||^2  ==  Kss - einsum(Kfs, solve(K + P^-1, Kfs)) (approximators.py:175-178)
            PB_TRY(gram_cross(st, prob->kernel, Zs, m, ws.Z(), n, Df, chunk, n, V, ld, ws.vec(V_S)));
            PB_TRY(trsm_right_lt(st, ws.B(), n, ld, ws.potrf_ws(), V, m, ld));
            row_sumsq_kernel<<<(unsigned)ceil_div<int64_t>(m, 8), 256, 0, st>>>(V, m, n, ld, kss, variance + r0); pb::note_launch();
            PB_CUDA(cudaGetLastError());
        }
    }
    return PB_OK;
}

// d objective / d (scale, stretch_out, sigma) at the converged weight: the closed form that replaces JAX's
// reverse pass through fixed_point_layer (probit/implicit/solvers.py:28-64, approximators.py:132-134).
// PRECONDITION: the workspace holds K(theta) and the Cholesky factor of B(w*) (pb_laplace_fit with
// final_factor = 1 was the last call).  The factor is consumed (the buffer ends holding B^-1).
// grad_host[0] = dPsi/dscale, [1] = dPsi/dstretch_out, [2] = dPsi/dsigma (NaN unless Gaussian).
extern "C" int64_t pb_gradient_scratch_bytes(int64_t n) {
    const int64_t ld = round_up(n > 0 ? n : 1, 16);
    const int64_t part = gram_deriv_partial_doubles(n) * 8;
    return n * ld * 8 > part ? n * ld * 8 : part;
}

extern "C" int pb_laplace_gradient(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes,
                                   const double* weight, const double* precision, void* scratch, int64_t scratch_bytes,
                                   double* grad_host, int32_t grad_len, const pb_options* options) {
    OptScope opt_scope(options);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    PB_CHECK(weight && precision && grad_host && grad_len >= 3, PB_ERR_INVALID, "laplace_gradient: bad argument");
    PB_CHECK(scratch && scratch_bytes >= pb_gradient_scratch_bytes(prob->n), PB_ERR_INVALID, "laplace_gradient: scratch too small");
    PB_CHECK((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, PB_ERR_INVALID, "laplace_gradient: scratch must be 256-byte aligned");
    lik::Params lp;
    PB_TRY(lik::make_params(prob->lik, lp));
    const int64_t n = prob->n, ld = ws.L.ld;
    const int Df = feature_dim(prob->kernel, prob->D);
    const unsigned nb = vec_blocks(n);
    const bool gaussian = prob->lik.kind == PB_LIK_GAUSSIAN;
    double* U = reinterpret_cast<double*>(scratch);
    double* sc = ws.scalars();
    double *f = ws.vec(V_F), *g = ws.vec(V_G), *d3 = ws.vec(V_U), *sv = ws.vec(V_S);

    PB_TRY(gemv(st, ws.K(), n, n, ld, weight, f));
    PB_TRY(likelihood(st, prob->lik, f, prob->y, n, 1, nullptr, g, nullptr, d3));
    sqrt_kernel<<<nb, 256, 0, st>>>(precision, n, curvature_floor(prob), sv, ws.partial()); pb::note_launch();
    // b_c * c = K g ;  b_l * l = (K o rho) g
    PB_TRY(gemv(st, ws.K(), n, n, ld, g, ws.vec(V_B)));
    PB_TRY(gram_deriv_matvec(st, prob->kernel, ws.Z(), n, Df, n, ws.K(), ld, g, ws.vec(V_T)));
    // s3 = b - K (s o B^-1 (s o b)) for both (still with the Cholesky factor in place)
    auto s3 = [&](double* b, double* out) -> int {
        mul_kernel<<<nb, 256, 0, st>>>(sv, b, n, ws.vec(V_C)); pb::note_launch();
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_C), ws.vec(V_X)));
        PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_X), ws.vec(V_C)));
        mul_kernel<<<nb, 256, 0, st>>>(sv, ws.vec(V_C), n, ws.vec(V_X)); pb::note_launch();
        PB_TRY(gemv(st, ws.K(), n, n, ld, ws.vec(V_X), ws.vec(V_Y)));
        sub_kernel<<<nb, 256, 0, st>>>(b, ws.vec(V_Y), n, out); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        return PB_OK;
    };
    PB_TRY(s3(ws.vec(V_B), ws.vec(V_R)));       // s3_c * c
    PB_TRY(s3(ws.vec(V_T), ws.vec(V_Z)));       // s3_l * l
    // U = L^-T (upper) in scratch; diag(B^-1) = row sums of squares of U -> V, s2 (the factor is still intact)
    PB_TRY(set_identity(st, U, n, ld));
    PB_TRY(trsm_right_lt(st, ws.B(), n, ld, ws.potrf_ws(), U, n, ld));
    row_sumsq_kernel<<<(unsigned)ceil_div<int64_t>(n, 8), 256, 0, st>>>(U, n, n, ld, 0.0, ws.vec(V_Q)); pb::note_launch();
    grad_s2_kernel<<<nb, 256, 0, st>>>(ws.vec(V_Q), precision, d3, f, prob->y, gaussian ? 1 : 0, prob->lik.sigma, n,
                                       curvature_floor(prob), prob->kernel.scale, ws.vec(V_P), ws.vec(V_E),
                                       ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_GSIG, sc + S_BAD));
    const bool ordinal_params = !gaussian && grad_len >= 3 + prob->lik.J + 1;
    if (ordinal_params) {
        // uvec = (K^-1 + W)^-1 s2 = K s2 - K R K s2 (needs the Cholesky factor: do it before B^-1 overwrites it)
        PB_TRY(gemv(st, ws.K(), n, n, ld, ws.vec(V_P), ws.vec(V_W)));
        PB_TRY(s3(ws.vec(V_W), ws.vec(V_WN)));
    }
    // B^-1 = U U^T (lower) over the factor
    PB_TRY(gemm_nt_mode(st, n, n, n, 1.0, U, ld, U, ld, 0.0, ws.B(), ld, 2));
    // lower-triangle sums (partials reuse the scratch: U is no longer needed)
    PB_TRY(gram_deriv_dots(st, prob->kernel, ws.Z(), n, Df, n, ws.K(), ld, ws.B(), ld, weight, sv, U, sc + S_G0));
    dot2_kernel<<<nb, 256, 0, st>>>(ws.vec(V_P), ws.vec(V_R), ws.vec(V_P), ws.vec(V_Z), n, ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_GD0, sc + S_GD1));
    dot2_kernel<<<nb, 256, 0, st>>>(f, weight, f, weight, n, ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_FTW, nullptr));
    double host[S_COUNT];
    PB_TRY(read_scalars(st, ws, host, nullptr));
    if (host[S_BAD] > 0) {
        set_error("laplace_gradient: %d precisions are materially negative or NaN", (int)host[S_BAD]);
        return PB_ERR_NUMERIC;
    }
    const double c = prob->kernel.scale, l = prob->kernel.stretch_out;
    const double dZ_c = (0.5 * host[S_FTW] - 0.5 * host[S_G2] + host[S_GD0]) / c;
    const double dZ_l = (0.5 * host[S_G0] - 0.5 * host[S_G1] + host[S_GD1]) / l;
    grad_host[0] = -dZ_c;
    grad_host[1] = -dZ_l;
    grad_host[2] = gaussian ? -host[S_GSIG] : NAN;
    if (ordinal_params) {
        const int J = prob->lik.J, width = J + 2;
        double* part = U;                                    // scratch is free again
        ordinal_param_grad_kernel<<<nb, 256, 0, st>>>(f, reinterpret_cast<const long long*>(prob->y), prob->lik.cutpoints, J,
                                                      prob->lik.sigma, prob->lik.eps, ws.vec(V_E), ws.vec(V_WN), n, part); pb::note_launch();
        sum_columns_kernel<<<1, 256, 0, st>>>(part, (int)nb, width, part + (int64_t)nb * width); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        std::vector<double> hostp(width);
        PB_CUDA(cudaMemcpyAsync(hostp.data(), part + (int64_t)nb * width, width * sizeof(double), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        grad_host[2] = -hostp[0];
        for (int j = 0; j <= J; ++j) grad_host[3 + j] = -hostp[1 + j];     // b_0 and b_J are infinite: their slots stay 0
    }
    return PB_OK;
}

// s = sqrt(-h), W = -h ; partial[1]: number of data with -h < 0 or NaN
__global__ void __launch_bounds__(256)
vb_curvature_kernel(const double* __restrict__ h, int64_t n, double neg_floor, double* __restrict__ s,
                    double* __restrict__ W, double* __restrict__ partial) {
    double bad = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double w = -h[i];
        if (!(w >= neg_floor)) bad += 1.0;
        w = w > 0.0 ? w : 0.0;
        W[i] = w;
        s[i] = sqrt(w);
    }
    write_partials(0.0, bad, partial);
}

// u = coef (f - kt), a = W o u
__global__ void __launch_bounds__(256)
vb_adjoint_kernel(const double* __restrict__ f, const double* __restrict__ kt, const double* __restrict__ W, double coef,
                  int64_t n, double* __restrict__ u, double* __restrict__ a) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double ui = coef * (f[i] - kt[i]);
        u[i] = ui;
        a[i] = W[i] * ui;
    }
}

// Gaussian likelihood: partial[0] = sum_i (dll_i/dsigma - sigma u_i dg_i/dsigma), r = y - f
__global__ void __launch_bounds__(256)
vb_gauss_sigma_kernel(const double* __restrict__ f, const double* __restrict__ y, const double* __restrict__ u,
                      double sigma, int64_t n, double* __restrict__ partial) {
    const double i3 = 1.0 / (sigma * sigma * sigma);
    double t = 0;
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const double r = y[i] - f[i];
        t += (-1.0 / sigma + r * r * i3) - sigma * u[i] * (-2.0 * r * i3);
    }
    write_partials(t, 0.0, partial);
}

// Gradient of objective_VB (VB.py:19-40) at the fixed point of f_VB (VB.py:4-16) with respect to the kernel's scale
// and outer stretch, the noise std and the cutpoints: what the reference obtains by reverse mode through
// fixed_point_layer (solvers.py:28-64, approximators.py:132-134,316-330), in the closed form derived and checked in
// oracle/gradients.py::vb_gradient:
//     dF/dphi = F_phi + u^T (M T_phi),   u = (1 - sigma)/sigma^2 (f - K S A^-1 S f),   A = sigma I + S K S,
//     M = sigma^2 I + K,  S = (-h)^1/2;  kernel parameter with C = dK/dphi:
//     dF/dphi = (1/2 - sigma) w^T C w + 1/2 tr(M^-1 C) - sigma (W u)^T C w.
// Two Cholesky factorisations (M for the traces, A for the adjoint solve), one N x N right-TRSM and one SYRK, all on
// the DMMA GEMM.  grad_host layout as pb_laplace_gradient: [d scale, d stretch_out, d sigma, d cutpoint_0..J].
extern "C" int pb_vb_gradient(pb_stream_t stream, const pb_problem* prob, void* workspace, int64_t workspace_bytes,
                              const double* weight, void* scratch, int64_t scratch_bytes, double* grad_host,
                              int32_t grad_len, const pb_options* options) {
    OptScope opt_scope(options);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Ws ws;
    PB_TRY(bind(prob, workspace, workspace_bytes, ws));
    PB_CHECK(weight && grad_host && grad_len >= 3, PB_ERR_INVALID, "vb_gradient: bad argument");
    PB_CHECK(scratch && scratch_bytes >= pb_gradient_scratch_bytes(prob->n), PB_ERR_INVALID, "vb_gradient: scratch too small");
    PB_CHECK((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, PB_ERR_INVALID, "vb_gradient: scratch must be 256-byte aligned");
    lik::Params lp;
    PB_TRY(lik::make_params(prob->lik, lp));
    const int64_t n = prob->n, ld = ws.L.ld;
    const int Df = feature_dim(prob->kernel, prob->D);
    const unsigned nb = vec_blocks(n);
    const bool gaussian = prob->lik.kind == PB_LIK_GAUSSIAN;
    const double sigma = prob->lik.sigma;
    double* U = reinterpret_cast<double*>(scratch);
    double* sc = ws.scalars();
    double *f = ws.vec(V_F), *sv = ws.vec(V_S), *Wv = ws.vec(V_Q), *ones = ws.vec(V_E), *u = ws.vec(V_P), *a = ws.vec(V_R);
    int32_t info_host = 0;
    double host[S_COUNT];

    PB_CUDA(cudaMemsetAsync(sc, 0, S_COUNT * sizeof(double), st));
    PB_TRY(gemv(st, ws.K(), n, n, ld, weight, f));
    PB_TRY(likelihood(st, prob->lik, f, prob->y, n, 1, nullptr, nullptr, ws.vec(V_T), nullptr));
    vb_curvature_kernel<<<nb, 256, 0, st>>>(ws.vec(V_T), n, curvature_floor(prob), sv, Wv, ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, nullptr, sc + S_BAD));
    fill_kernel<<<nb, 256, 0, st>>>(ones, n, 1.0); pb::note_launch();
    // M = sigma^2 I + K: U = L^-T, diag(M^-1) = row sums of squares of U, M^-1 = U U^T over the factor
    PB_TRY(factor_matrix(st, ws, n, nullptr, sigma * sigma, 0.0));
    PB_TRY(set_identity(st, U, n, ld));
    PB_TRY(trsm_right_lt(st, ws.B(), n, ld, ws.potrf_ws(), U, n, ld));
    row_sumsq_kernel<<<(unsigned)ceil_div<int64_t>(n, 8), 256, 0, st>>>(U, n, n, ld, 0.0, ws.vec(V_X)); pb::note_launch();
    dot2_kernel<<<nb, 256, 0, st>>>(ws.vec(V_X), ones, f, weight, n, ws.partial()); pb::note_launch();    // -tr(M^-1), f.w
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_GSIG, sc + S_FTW));
    PB_TRY(gemm_nt_mode(st, n, n, n, 1.0, U, ld, U, ld, 0.0, ws.B(), ld, 2));
    // S_G0 = w^T (K o rho) w, S_G1 = tr(M^-1 (K o rho)), S_G2 = tr(M^-1 K)   (K o rho = l dK/dl, K = c dK/dc)
    PB_TRY(gram_deriv_dots(st, prob->kernel, ws.Z(), n, Df, n, ws.K(), ld, ws.B(), ld, weight, ones, U, sc + S_G0));
    PB_TRY(read_scalars(st, ws, host, &info_host));
    if (info_host != 0 || host[S_BAD] > 0) {
        set_error("vb_gradient: sigma^2 I + K is not positive definite (info %d) or %d data have negative curvature",
                  info_host, (int)host[S_BAD]);
        return PB_ERR_NUMERIC;
    }
    // A = sigma I + S K S: u = (1 - sigma)/sigma^2 (f - K S A^-1 S f), a = W o u
    PB_TRY(factor_matrix(st, ws, n, sv, sigma, 0.0));
    mul_kernel<<<nb, 256, 0, st>>>(sv, f, n, ws.vec(V_C)); pb::note_launch();
    PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), false, ws.vec(V_C), ws.vec(V_X)));
    PB_TRY(trsv(st, ws.B(), n, ld, ws.dinv(), true, ws.vec(V_X), ws.vec(V_C)));
    mul_kernel<<<nb, 256, 0, st>>>(sv, ws.vec(V_C), n, ws.vec(V_X)); pb::note_launch();
    PB_TRY(gemv(st, ws.K(), n, n, ld, ws.vec(V_X), ws.vec(V_Y)));
    vb_adjoint_kernel<<<nb, 256, 0, st>>>(f, ws.vec(V_Y), Wv, (1.0 - sigma) / (sigma * sigma), n, u, a); pb::note_launch();
    PB_TRY(gram_deriv_matvec(st, prob->kernel, ws.Z(), n, Df, n, ws.K(), ld, weight, ws.vec(V_Z)));       // (K o rho) w
    dot2_kernel<<<nb, 256, 0, st>>>(a, f, a, ws.vec(V_Z), n, ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_GD0, sc + S_GD1));
    dot2_kernel<<<nb, 256, 0, st>>>(u, weight, u, weight, n, ws.partial()); pb::note_launch();
    PB_CUDA(cudaGetLastError());
    PB_TRY(finalize(st, ws, nb, sc + S_RZ0, nullptr));
    double t_sigma = 0.0;
    const int J = prob->lik.J, width = J + 2;
    std::vector<double> hostp(gaussian ? 1 : width, 0.0);
    if (gaussian) {
        vb_gauss_sigma_kernel<<<nb, 256, 0, st>>>(f, reinterpret_cast<const double*>(prob->y), u, sigma, n, ws.partial()); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_TRY(finalize(st, ws, nb, sc + S_RZ1, nullptr));
    } else {
        // per datum dll/dphi - sigma u dg/dphi: the Laplace kernel with V = 0 and uvec = -sigma u
        double* part = U;
        fill_kernel<<<nb, 256, 0, st>>>(ones, n, 0.0); pb::note_launch();
        vb_adjoint_kernel<<<nb, 256, 0, st>>>(u, ones, ones, -sigma, n, ws.vec(V_WN), ws.vec(V_X)); pb::note_launch();  // -sigma (u - 0)
        ordinal_param_grad_kernel<<<nb, 256, 0, st>>>(f, reinterpret_cast<const long long*>(prob->y), prob->lik.cutpoints, J,
                                                      sigma, prob->lik.eps, ones, ws.vec(V_WN), n, part); pb::note_launch();
        sum_columns_kernel<<<1, 256, 0, st>>>(part, (int)nb, width, part + (int64_t)nb * width); pb::note_launch();
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaMemcpyAsync(hostp.data(), part + (int64_t)nb * width, width * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    PB_TRY(read_scalars(st, ws, host, &info_host));
    if (info_host != 0) {
        set_error("vb_gradient: sigma I + W^1/2 K W^1/2 is not positive definite (info %d)", info_host);
        return PB_ERR_NUMERIC;
    }
    t_sigma = gaussian ? host[S_RZ1] : hostp[0];
    const double c = prob->kernel.scale, l = prob->kernel.stretch_out;
    const double tr_minv = -host[S_GSIG];
    grad_host[0] = ((0.5 - sigma) * host[S_FTW] + 0.5 * host[S_G2] - sigma * host[S_GD0]) / c;
    grad_host[1] = ((0.5 - sigma) * host[S_G0] + 0.5 * host[S_G1] - sigma * host[S_GD1]) / l;
    grad_host[2] = -t_sigma - (double)n / sigma + sigma * tr_minv - sigma * host[S_RZ0];
    if (!gaussian)
        for (int j = 0; j <= J && 3 + j < grad_len; ++j) grad_host[3 + j] = -hostp[1 + j];
    return PB_OK;
}

// predict_covariance (probit/approximators.py:182-197): C = K_** - K_*f (K + P^-1)^-1 K_f* for n_test points
// = K_** - V V^T with V = (s o K_*f) L_B^-T.  Needs the factor from pb_predict_prepare.  `scratch` holds V
// (pb_predict_scratch_bytes(n, D, n_test)); `cov` is n_test x ldc, full (both triangles) on return.
extern "C" int pb_predict_covariance(pb_stream_t stream, const pb_problem* prob, const void* workspace,
                                     const double* X_test, int64_t n_test, void* scratch, int64_t scratch_bytes,
                                     double* cov, int64_t ldc) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    PB_TRY(check_problem(prob));
    PB_CHECK(workspace && X_test && cov && n_test >= 1, PB_ERR_INVALID, "predict_covariance: bad argument");
    PB_CHECK(scratch && scratch_bytes >= pb_predict_scratch_bytes(prob->n, prob->D, n_test), PB_ERR_INVALID,
             "predict_covariance: scratch too small");
    Ws ws;
    ws.L = make_layout(prob->n, prob->D);
    ws.base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(workspace));
    const int64_t n = prob->n, ld = ws.L.ld;
    const int D = prob->D, Df = feature_dim(prob->kernel, D);
    double* V = reinterpret_cast<double*>(scratch);
    double* Zs = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(scratch) + round_up(n_test * ld * 8, 256));
    PB_TRY(features(st, prob->kernel, X_test, n_test, D, D, Zs, n_test));
    PB_TRY(gram_sym(st, prob->kernel, Zs, n_test, Df, n_test, cov, ldc, nullptr, 0.0));                  // K_**
    PB_TRY(gram_cross(st, prob->kernel, Zs, n_test, ws.Z(), n, Df, n_test, n, V, ld, ws.vec(V_S)));       // s o K_*f
    PB_TRY(trsm_right_lt(st, ws.B(), n, ld, ws.potrf_ws(), V, n_test, ld));
    return gemm_nt(st, n_test, n_test, n, -1.0, V, ld, V, ld, 1.0, cov, ldc, false);
}
