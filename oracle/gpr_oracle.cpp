// gpr_oracle.cpp — CPU restatement of the reference GP-regression hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (include/, the CUDA
// library, the C++ drop-in headers) links, loads or calls this file.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use it, and only as the checker or the timed CPU baseline.
//
// Parity status: the reference's own tests pin no numerical result and
// Eigen, where the reference's dense arithmetic lives, is absent from this image
// (module Eigen3, no pinned version: reference CMakeLists.txt:25).
// The restatement is therefore pinned in two ways (see oracle/README.md):
//   1. against the reference's OWN header include/gp_regression/gp_regressor.hpp
//      compiled unmodified from /root/reference against the API shim in
//      oracle/eigen_shim (oracle/ref_driver.cpp -> oracle/_ref/ref_driver), and
//   2. against closed-form known answers.
// The LDLT arithmetic itself is restated from Eigen's published unblocked
// algorithm (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked) in both
// places, so the factorisation is "restated", not "pinned".
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/include/gp_regression unless stated).
//
// Build: g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -pthread

#include <immintrin.h>
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------------
// Covariance functors.  kernels/thin_plate.hpp:12-20, kernels/gaussian.hpp:15-27,
// kernels/laplace.hpp:37-49.  kind: 0 ThinPlate(R=p0), 1 Gaussian(sigma=p0,length=p1),
// 2 Laplace(sigma=p0,length=p1).  "diff" is the reference's computediff, which is
// (1/d)*dk/dd for ThinPlate but dk/dd for Gaussian/Laplace — reproduced.
// ---------------------------------------------------------------------------------
template <typename T>
struct Kern {
    int kind;
    T p0, p1;
    T R3, sigma2, inv_length2, inv_length;
    Kern(int k, double a, double b) : kind(k), p0((T)a), p1((T)b) {
        R3 = p0 * p0 * p0;                    // thin_plate.hpp:31
        sigma2 = p0 * p0;                     // gaussian.hpp:40
        inv_length2 = (T)1.0 / (p1 * p1);     // gaussian.hpp:41
        inv_length = (T)1.0 / p1;             // laplace.hpp:62
    }
    T compute(T d) const {
        switch (kind) {
            case 0: return 2 * d * d * d - 3 * p0 * d * d + R3;           // thin_plate.hpp:14
            case 1: return sigma2 * std::exp(-1 * d * inv_length2);       // gaussian.hpp:17-18
            default: return 2 * p0 * std::exp(-1 * d * inv_length);      // laplace.hpp:39-40
        }
    }
    T diff(T d) const {
        switch (kind) {
            case 0: return -6 * (p0 - d);                                 // thin_plate.hpp:19
            case 1: return -1 * inv_length2 * compute(d);                 // gaussian.hpp:24-25
            default: return -1 * inv_length * compute(d);                // laplace.hpp:46-47
        }
    }
};

// ---------------------------------------------------------------------------------
// Pairwise distance.  gp_regressor.hpp:548-557 uses sqrt(|a|^2+|b|^2-2a.b) with no
// clamp.  dist_mode 0 = difference form sqrt(sum (a-b)^2) (documented deviation (i),
// SURVEY §8c; what the CUDA kernels compute); dist_mode 1 = the reference's expansion,
// evaluated in the reference's operation order ((-2*a).b accumulated over x,y,z, then
// += |a|^2, then += |b|^2), so that the diagonal cancels exactly when built without
// FMA contraction, as on the reference's default x86-64 build.
// ---------------------------------------------------------------------------------
template <typename T>
static inline T dist(const T* a, const T* b, int mode) {
    if (mode == 0) {
        T dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    T m2ab = ((-2 * a[0]) * b[0] + (-2 * a[1]) * b[1]) + (-2 * a[2]) * b[2];
    T aa = (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2];
    T bb = (b[0] * b[0] + b[1] * b[1]) + b[2] * b[2];
    return std::sqrt((m2ab + aa) + bb);
}

// ---------------------------------------------------------------------------------
// Factorisations of a dense symmetric matrix held column-major in A (n x n, lower
// triangle referenced).
// ---------------------------------------------------------------------------------

// Plain lower Cholesky (LLT), right-looking by columns.  Returns 0 or 1+index of the
// first non-positive pivot.  This is what the CUDA path computes.
template <typename T>
static int llt_inplace(std::vector<T>& A, int n) {
    for (int k = 0; k < n; ++k) {
        T* ck = &A[(size_t)k * n];
        T akk = ck[k];
        if (!(akk > 0)) return k + 1;
        T l = std::sqrt(akk);
        ck[k] = l;
        T inv = (T)1 / l;
        for (int i = k + 1; i < n; ++i) ck[i] *= inv;
        for (int j = k + 1; j < n; ++j) {
            T ljk = ck[j];
            if (ljk == 0) continue;
            T* cj = &A[(size_t)j * n];
            for (int i = j; i < n; ++i) cj[i] -= ck[i] * ljk;
        }
    }
    return 0;
}

// Diagonal-pivoted LDL^T following Eigen's unblocked algorithm
// (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked, Eigen 3.3 line of code;
// called from gp_regressor.hpp:161-162 and :457-458).  On exit the strict lower
// triangle of A holds L (unit diagonal implied), the diagonal holds D, and
// perm[k] = index swapped with k at step k (Eigen's "transpositions").
template <typename T>
static void ldlt_inplace(std::vector<T>& A, int n, std::vector<int>& perm) {
    perm.assign(n, 0);
    std::vector<T> temp(n);
    auto at = [&](int r, int c) -> T& { return A[(size_t)c * n + r]; };
    if (n <= 1) {
        if (n == 1) perm[0] = 0;
        return;
    }
    for (int k = 0; k < n; ++k) {
        // biggest |diagonal| entry in the remaining corner
        int piv = k;
        T best = std::fabs(at(k, k));
        for (int i = k + 1; i < n; ++i) {
            T v = std::fabs(at(i, i));
            if (v > best) { best = v; piv = i; }
        }
        perm[k] = piv;
        if (piv != k) {
            // symmetric row/column interchange touching the lower triangle only
            int s = n - piv - 1;
            for (int c = 0; c < k; ++c) std::swap(at(k, c), at(piv, c));
            for (int i = 0; i < s; ++i) std::swap(at(piv + 1 + i, k), at(piv + 1 + i, piv));
            std::swap(at(k, k), at(piv, piv));
            for (int i = k + 1; i < piv; ++i) std::swap(at(i, k), at(piv, i));
        }
        // A10 = row k, cols [0,k);  A20 = rows (k,n), cols [0,k);  A21 = rows (k,n), col k
        if (k > 0) {
            for (int c = 0; c < k; ++c) temp[c] = at(c, c) * at(k, c);
            T dot = 0;
            for (int c = 0; c < k; ++c) dot += at(k, c) * temp[c];
            at(k, k) -= dot;
            for (int c = 0; c < k; ++c) {
                T t = temp[c];
                if (t == 0) continue;
                const T* col = &A[(size_t)c * n];
                T* dst = &A[(size_t)k * n];
                for (int i = k + 1; i < n; ++i) dst[i] -= col[i] * t;
            }
        }
        T akk = at(k, k);
        bool valid = std::fabs(akk) > 0;
        if (k == 0 && !valid) {
            for (int j = 0; j < n; ++j) perm[j] = j;
            return;
        }
        if (valid) {
            T* dst = &A[(size_t)k * n];
            for (int i = k + 1; i < n; ++i) dst[i] /= akk;
        }
    }
}

// x <- K^{-1} x using the LDLT above (Eigen LDLT::_solve_impl: P, L^-1, D^-1 with the
// 1/highest tolerance of Eigen >= 3.3, L^-T, P^T).  gp_regressor.hpp:163, :263, :316.
template <typename T>
static void ldlt_solve(const std::vector<T>& A, int n, const std::vector<int>& perm, T* x) {
    auto at = [&](int r, int c) -> const T& { return A[(size_t)c * n + r]; };
    for (int k = 0; k < n; ++k) std::swap(x[k], x[perm[k]]);
    for (int c = 0; c < n; ++c) {
        T xc = x[c];
        if (xc == 0) continue;
        const T* col = &A[(size_t)c * n];
        for (int i = c + 1; i < n; ++i) x[i] -= col[i] * xc;
    }
    const T tol = std::numeric_limits<T>::min();
    for (int i = 0; i < n; ++i) {
        T d = at(i, i);
        if (std::fabs(d) > tol) x[i] /= d; else x[i] = 0;
    }
    for (int c = n - 1; c >= 0; --c) {
        const T* col = &A[(size_t)c * n];
        T s = x[c];
        for (int i = c + 1; i < n; ++i) s -= col[i] * x[i];
        x[c] = s;
    }
    for (int k = n - 1; k >= 0; --k) std::swap(x[k], x[perm[k]]);
}

template <typename T>
static void llt_solve(const std::vector<T>& A, int n, T* x) {
    for (int c = 0; c < n; ++c) {
        const T* col = &A[(size_t)c * n];
        x[c] /= col[c];
        T xc = x[c];
        for (int i = c + 1; i < n; ++i) x[i] -= col[i] * xc;
    }
    for (int c = n - 1; c >= 0; --c) {
        const T* col = &A[(size_t)c * n];
        T s = x[c];
        for (int i = c + 1; i < n; ++i) s -= col[i] * x[i];
        x[c] = s / col[c];
    }
}

// ---------------------------------------------------------------------------------
// Model: the subset of gp_regression::Model (gp_regressor.hpp:71-87) the path uses.
// ---------------------------------------------------------------------------------
template <typename T>
struct Model {
    int n = 0;
    int kind = 0;
    double p0 = 1, p1 = 1;
    int factor_mode = 0;  // 0 = pivoted LDLT (reference), 1 = LLT
    int dist_mode = 0;
    bool has_sigma2 = false;
    T R = 0;
    std::vector<T> P;      // n x 3 row-major here (x,y,z per point)
    std::vector<T> Y, S2, alpha;
    std::vector<T> Kpp;    // n x n column-major, as assembled (before factorisation)
    std::vector<T> F;      // factor (LDLT or LLT), column-major
    std::vector<int> perm;
    std::vector<T> N;      // n x 3 column-major normals (withNormals)
    int info = 0;          // LLT: 1+index of failing pivot; LDLT: 0
};

template <typename T>
static void factor_and_solve(Model<T>& m) {
    int n = m.n;
    m.F = m.Kpp;
    m.info = 0;
    if (m.factor_mode == 0) ldlt_inplace(m.F, n, m.perm);
    else m.info = llt_inplace(m.F, n);
    m.alpha = m.Y;
    if (m.info == 0) {
        if (m.factor_mode == 0) ldlt_solve(m.F, n, m.perm, m.alpha.data());
        else llt_solve(m.F, n, m.alpha.data());
    }
}

// create<withNormals>: gp_regressor.hpp:110-182.
template <typename T>
static void fit(Model<T>& m, const double* x, const double* y, const double* z,
                const double* label, const double* sigma2, int n, bool with_normals) {
    m.n = n;
    m.P.resize((size_t)n * 3);
    m.Y.resize(n);
    m.has_sigma2 = sigma2 != nullptr;
    m.S2.assign(n, 0);
    for (int i = 0; i < n; ++i) {                          // :120-122
        m.P[3 * i + 0] = (T)x[i]; m.P[3 * i + 1] = (T)y[i]; m.P[3 * i + 2] = (T)z[i];
        m.Y[i] = (T)label[i];
        if (sigma2) m.S2[i] = (T)sigma2[i];
    }
    Kern<T> kern(m.kind, m.p0, m.p1);
    m.Kpp.assign((size_t)n * n, 0);
    std::vector<T> Kdiff;
    if (with_normals) Kdiff.assign((size_t)n * n, 0);      // :140
    T R = 0;
    // :132 distance matrix, :135 R = max, :144-159 kernel map with sigma2 on the diagonal.
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            T d = dist(&m.P[3 * i], &m.P[3 * j], m.dist_mode);
            if (d > R) R = d;
            if (with_normals) Kdiff[(size_t)j * n + i] = kern.diff(d);   // :152
            T k = kern.compute(d);
            if (m.has_sigma2 && i == j) k += m.S2[i];                    // :154-155
            m.Kpp[(size_t)j * n + i] = k;
        }
    m.R = R;
    factor_and_solve(m);                                   // :161-163
    m.N.clear();
    if (with_normals && m.info == 0) {                     // :166-181 (zero-initialised: deviation (ii))
        m.N.assign((size_t)n * 3, 0);
        for (int i = 0; i < n; ++i) {
            T g[3] = {0, 0, 0};
            for (int j = 0; j < n; ++j) {
                T w = m.alpha[j] * Kdiff[(size_t)j * n + i];             // :172
                for (int c = 0; c < 3; ++c) g[c] += w * (m.P[3 * i + c] - m.P[3 * j + c]);
            }
            T nrm = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);  // :174 (Eigen normalize: no-op if norm==0)
            for (int c = 0; c < 3; ++c) m.N[(size_t)c * n + i] = nrm > 0 ? g[c] / nrm : g[c];
        }
    }
}

// evaluate (all overloads): gp_regressor.hpp:222-273 (f,v,N), :282-324 (f,v), :332-357 (f).
// Variance restated as the diagonal only, k(0) - k*^T K^-1 k* (deviation (iii));
// the reference takes the diagonal of Kqq - Kqp*solve(Kpq) (:263-266) where Kqq_ii = k(D_ii).
// grad is q x 3 column-major, un-normalised (:247-250).
template <typename T>
static void predict_range(const Model<T>& m, const double* qx, const double* qy, const double* qz,
                          int q0, int q1, int q, double* f, double* v, double* grad) {
    Kern<T> kern(m.kind, m.p0, m.p1);
    int n = m.n;
    std::vector<T> ks(n);
    for (int i = q0; i < q1; ++i) {
        T qq[3] = {(T)qx[i], (T)qy[i], (T)qz[i]};
        T g[3] = {0, 0, 0};
        T acc = 0;
        for (int j = 0; j < n; ++j) {
            T d = dist(qq, &m.P[3 * j], m.dist_mode);
            if (grad) {
                T w = m.alpha[j] * kern.diff(d);                          // :247
                for (int c = 0; c < 3; ++c) g[c] += w * (qq[c] - m.P[3 * j + c]);
            }
            T k = kern.compute(d);                                        // :248, :303, :351
            ks[j] = k;
            acc += k * m.alpha[j];                                        // :252, :305, :353
        }
        f[i] = (double)acc;
        if (grad) for (int c = 0; c < 3; ++c) grad[(size_t)c * q + i] = (double)g[c];
        if (v) {
            std::vector<T> w(ks);
            if (m.factor_mode == 0) ldlt_solve(m.F, n, m.perm, w.data());  // :263, :316
            else llt_solve(m.F, n, w.data());
            T s = 0;
            for (int j = 0; j < n; ++j) s += ks[j] * w[j];                // :265, :318
            T dqq = dist(qq, qq, m.dist_mode);                            // :256, :309 (0 in both forms)
            v[i] = (double)(kern.compute(dqq) - s);                       // :266, :319
        }
    }
}

template <typename T>
static void predict(const Model<T>& m, const double* qx, const double* qy, const double* qz, int q,
                    double* f, double* v, double* grad, int threads) {
    if (threads <= 1 || q < 2) { predict_range(m, qx, qy, qz, 0, q, q, f, v, grad); return; }
    // One host thread per slice of queries: mirrors how the node fans evaluate() out over
    // std::threads (reference src/gp_node.cpp:1027-1038).
    std::vector<std::thread> pool;
    int t = std::min(threads, q);
    for (int r = 0; r < t; ++r) {
        int a = (int)((long long)q * r / t), b = (int)((long long)q * (r + 1) / t);
        pool.emplace_back([&, a, b] { predict_range(m, qx, qy, qz, a, b, q, f, v, grad); });
    }
    for (auto& th : pool) th.join();
}

// update<withNormals>: gp_regressor.hpp:367-479 — append, then refactorise from scratch
// (SURVEY F6).  R and normals are not refreshed (:454-455, :462-477).
template <typename T>
static void update(Model<T>& m, const double* x, const double* y, const double* z,
                   const double* label, const double* sigma2, int k) {
    int p = m.n, n = p + k;
    Kern<T> kern(m.kind, m.p0, m.p1);
    std::vector<T> P((size_t)n * 3), Y(n), S2(n, 0), K((size_t)n * n, 0);
    std::copy(m.P.begin(), m.P.end(), P.begin());
    std::copy(m.Y.begin(), m.Y.end(), Y.begin());
    std::copy(m.S2.begin(), m.S2.end(), S2.begin());
    for (int i = 0; i < k; ++i) {
        P[3 * (p + i) + 0] = (T)x[i]; P[3 * (p + i) + 1] = (T)y[i]; P[3 * (p + i) + 2] = (T)z[i];
        Y[p + i] = (T)label[i];
        if (sigma2) S2[p + i] = (T)sigma2[i];
    }
    for (int j = 0; j < p; ++j)                                             // :442 conservativeResize keeps old block
        for (int i = 0; i < p; ++i) K[(size_t)j * n + i] = m.Kpp[(size_t)j * p + i];
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < p; ++i) {                                       // :398, :408-421, :444-445
            T v = kern.compute(dist(&P[3 * i], &P[3 * (p + j)], m.dist_mode));
            K[(size_t)(p + j) * n + i] = v;
            K[(size_t)i * n + (p + j)] = v;
        }
    for (int j = 0; j < k; ++j)
        for (int i = 0; i < k; ++i) {                                       // :397, :424-440, :443
            T v = kern.compute(dist(&P[3 * (p + i)], &P[3 * (p + j)], m.dist_mode));
            if (sigma2 && i == j) v += (T)sigma2[i];                        // :435-436
            K[(size_t)(p + j) * n + (p + i)] = v;
        }
    m.n = n; m.P.swap(P); m.Y.swap(Y); m.S2.swap(S2); m.Kpp.swap(K);
    factor_and_solve(m);                                                    // :457-459
}

// computeTangentBasis: gp_regressor.hpp:29-44.  isApprox(a,b,p) is |a-b|^2 <= p^2 min(|a|^2,|b|^2).
template <typename T>
static void tangent_basis(const T g[3], T N[3], T Tx[3], T Ty[3]) {
    T nrm = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    for (int c = 0; c < 3; ++c) N[c] = nrm > 0 ? g[c] / nrm : g[c];          // :31
    T dx = N[0] - 1, nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    T d2 = dx * dx + N[1] * N[1] + N[2] * N[2];
    bool approx_x = d2 <= (T)1e-6 * std::min(nn, (T)1);                       // :32
    T e[3] = {approx_x ? (T)0 : (T)1, approx_x ? (T)1 : (T)0, 0};             // :33 / :39
    T ne = N[0] * e[0] + N[1] * e[1] + N[2] * e[2];
    for (int c = 0; c < 3; ++c) Tx[c] = e[c] - N[c] * ne;
    T tn = std::sqrt(Tx[0] * Tx[0] + Tx[1] * Tx[1] + Tx[2] * Tx[2]);
    if (tn > 0) for (int c = 0; c < 3; ++c) Tx[c] /= tn;                      // :34 / :40
    Ty[0] = N[1] * Tx[2] - N[2] * Tx[1];                                      // :35 / :41
    Ty[1] = N[2] * Tx[0] - N[0] * Tx[2];
    Ty[2] = N[0] * Tx[1] - N[1] * Tx[0];
    T yn = std::sqrt(Ty[0] * Ty[0] + Ty[1] * Ty[1] + Ty[2] * Ty[2]);
    if (yn > 0) for (int c = 0; c < 3; ++c) Ty[c] /= yn;                      // :36 / :42
}

}  // namespace orc

// hi/lo (rows x m4, double-double accumulators) += kt (rows x jn, pitch ktp) * Wt (jn x m4), every product exact
// (p = k*w, e = fma(k, w, -p)) and every sum a Knuth two-sum; 4 columns per AVX2 lane group, 16 columns of
// accumulators held in registers across the jn loop.
namespace orc {
__attribute__((target("avx2,fma")))
static void dd_accumulate_avx2(const double* kt, int ktp, int rows, int jn, const double* Wt, int m4, double* hi, double* lo) {
    for (int r = 0; r < rows; ++r) {
        const double* kr = kt + (size_t)r * ktp;
        for (int c0 = 0; c0 < m4; c0 += 16) {
            const int nv = std::min(4, (m4 - c0) / 4);
            __m256d h[4], l[4];
            for (int v = 0; v < nv; ++v) { h[v] = _mm256_loadu_pd(hi + (size_t)r * m4 + c0 + 4 * v); l[v] = _mm256_loadu_pd(lo + (size_t)r * m4 + c0 + 4 * v); }
            for (int jj = 0; jj < jn; ++jj) {
                const __m256d k = _mm256_set1_pd(kr[jj]);
                const double* w = Wt + (size_t)jj * m4 + c0;
                for (int v = 0; v < nv; ++v) {
                    const __m256d wv = _mm256_loadu_pd(w + 4 * v);
                    const __m256d p = _mm256_mul_pd(k, wv);
                    const __m256d e = _mm256_fmsub_pd(k, wv, p);            // exact error of the product
                    const __m256d s = _mm256_add_pd(h[v], p);               // two-sum(h, p)
                    const __m256d bb = _mm256_sub_pd(s, h[v]);
                    const __m256d err = _mm256_add_pd(_mm256_sub_pd(h[v], _mm256_sub_pd(s, bb)), _mm256_sub_pd(p, bb));
                    h[v] = s;
                    l[v] = _mm256_add_pd(l[v], _mm256_add_pd(err, e));
                }
            }
            for (int v = 0; v < nv; ++v) { _mm256_storeu_pd(hi + (size_t)r * m4 + c0 + 4 * v, h[v]); _mm256_storeu_pd(lo + (size_t)r * m4 + c0 + 4 * v, l[v]); }
        }
    }
}
}  // namespace orc

// ---------------------------------------------------------------------------------
// C interface for ctypes.  precision: 0 = double, 1 = long double (x87 80-bit; used
// to attribute error between two double implementations).
// ---------------------------------------------------------------------------------
struct orc_model {
    int precision;
    orc::Model<double> d;
    orc::Model<long double> l;
};

template <typename T>
static void copy_out(const std::vector<T>& src, double* dst) {
    if (!dst) return;
    for (size_t i = 0; i < src.size(); ++i) dst[i] = (double)src[i];
}

extern "C" {

orc_model* orc_fit(const double* x, const double* y, const double* z, const double* label,
                   const double* sigma2_or_null, int n, int kind, double p0, double p1,
                   int factor_mode, int dist_mode, int with_normals, int precision) {
    orc_model* m = new orc_model();
    m->precision = precision;
    if (precision == 0) {
        m->d.kind = kind; m->d.p0 = p0; m->d.p1 = p1; m->d.factor_mode = factor_mode; m->d.dist_mode = dist_mode;
        orc::fit(m->d, x, y, z, label, sigma2_or_null, n, with_normals != 0);
    } else {
        m->l.kind = kind; m->l.p0 = p0; m->l.p1 = p1; m->l.factor_mode = factor_mode; m->l.dist_mode = dist_mode;
        orc::fit(m->l, x, y, z, label, sigma2_or_null, n, with_normals != 0);
    }
    return m;
}

void orc_free(orc_model* m) { delete m; }

int orc_n(const orc_model* m) { return m->precision == 0 ? m->d.n : m->l.n; }
int orc_info(const orc_model* m) { return m->precision == 0 ? m->d.info : m->l.info; }
double orc_R(const orc_model* m) { return m->precision == 0 ? m->d.R : (double)m->l.R; }

// alpha[n]; normals n x 3 column-major or null; K n x n column-major or null; factor or null.
void orc_get(const orc_model* m, double* alpha, double* normals, double* K, double* factor) {
    if (m->precision == 0) {
        copy_out(m->d.alpha, alpha); if (!m->d.N.empty()) copy_out(m->d.N, normals);
        copy_out(m->d.Kpp, K); copy_out(m->d.F, factor);
    } else {
        copy_out(m->l.alpha, alpha); if (!m->l.N.empty()) copy_out(m->l.N, normals);
        copy_out(m->l.Kpp, K); copy_out(m->l.F, factor);
    }
}

void orc_predict(const orc_model* m, const double* qx, const double* qy, const double* qz, int q,
                 double* f, double* v_or_null, double* grad_or_null, int threads) {
    if (m->precision == 0) orc::predict(m->d, qx, qy, qz, q, f, v_or_null, grad_or_null, threads);
    else orc::predict(m->l, qx, qy, qz, q, f, v_or_null, grad_or_null, threads);
}

void orc_update(orc_model* m, const double* x, const double* y, const double* z, const double* label,
                const double* sigma2_or_null, int k) {
    if (m->precision == 0) orc::update(m->d, x, y, z, label, sigma2_or_null, k);
    else orc::update(m->l, x, y, z, label, sigma2_or_null, k);
}

// grad, N, Tx, Ty are q x 3 column-major.  gp_regressor.hpp:194-212.
void orc_tangent_basis(const double* grad, int q, double* N, double* Tx, double* Ty) {
    for (int i = 0; i < q; ++i) {
        double g[3] = {grad[i], grad[(size_t)q + i], grad[(size_t)2 * q + i]}, n[3], tx[3], ty[3];
        orc::tangent_basis(g, n, tx, ty);
        for (int c = 0; c < 3; ++c) {
            N[(size_t)c * q + i] = n[c]; Tx[(size_t)c * q + i] = tx[c]; Ty[(size_t)c * q + i] = ty[c];
        }
    }
}

// Extended-precision residuals R = B - K W for m right-hand sides (B, W, R: n x m column-major), K assembled entry by
// entry as create() does (gp_regressor.hpp:144-159: k(D_ij) + [i == j] sigma2_i; difference-form distance, the
// documented deviation (i)) and never stored.  Every product K_ij * W_jc is formed exactly (FMA error term) and
// accumulated in double-double (Knuth two-sum), i.e. with ~106-bit sums; without AVX2+FMA hardware the fallback
// accumulates in long double (x87, 64-bit mantissa) in blocks of 256.
// Used by the headline-size parity tests (n = 16 384 / 65 536), where no CPU factorisation fits a test's time budget:
// for ANY approximate solution w of K w = k*, v = k(0) - k*^T w - w^T r - O(|r|^2 / lambda_min(K)) with r = k* - K w,
// so a small certified residual pins the variance independently of how w was obtained.
void orc_residual(const double* x, const double* y, const double* z, const double* sigma2_or_null, int n, int kind,
                  double p0, double p1, const double* B, const double* W, int m, double* R, int threads) {
    orc::Kern<double> kern(kind, p0, p1);
    const int m4 = (m + 3) / 4 * 4;
    std::vector<double> Wt((size_t)n * m4, 0.0);            // row-major copy, rows padded to 4: Wt[j*m4 + c]
    for (int c = 0; c < m; ++c)
        for (int j = 0; j < n; ++j) Wt[(size_t)j * m4 + c] = W[(size_t)c * n + j];
    if (threads < 1) threads = 1;
    const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    constexpr int RT = 16, JB = 256;                        // row tile, column block
    const int ntiles = (n + RT - 1) / RT;
    auto work = [&](int t) {
        std::vector<double> hi((size_t)RT * m4), lo((size_t)RT * m4), kt((size_t)RT * JB);
        std::vector<long double> tot((size_t)RT * m4);
        for (int tile = t; tile < ntiles; tile += threads) {
            const int i0 = tile * RT, rows = std::min(RT, n - i0);
            std::fill(hi.begin(), hi.end(), 0.0); std::fill(lo.begin(), lo.end(), 0.0);
            std::fill(tot.begin(), tot.end(), 0.0L);
            for (int j0 = 0; j0 < n; j0 += JB) {
                const int jn = std::min(JB, n - j0);
                for (int r = 0; r < rows; ++r) {
                    const int i = i0 + r;
                    const double pi[3] = {x[i], y[i], z[i]};
                    for (int jj = 0; jj < jn; ++jj) {
                        const int j = j0 + jj;
                        const double pj[3] = {x[j], y[j], z[j]};
                        double kij = kern.compute(orc::dist(pi, pj, 0));
                        if (i == j && sigma2_or_null) kij += sigma2_or_null[i];
                        kt[(size_t)r * JB + jj] = kij;
                    }
                }
                if (fast) orc::dd_accumulate_avx2(kt.data(), JB, rows, jn, &Wt[(size_t)j0 * m4], m4, hi.data(), lo.data());
                else {
                    std::vector<long double> blk(m4);
                    for (int r = 0; r < rows; ++r) {
                        std::fill(blk.begin(), blk.end(), 0.0L);
                        for (int jj = 0; jj < jn; ++jj) {
                            const long double kl = kt[(size_t)r * JB + jj];
                            const double* wr = &Wt[(size_t)(j0 + jj) * m4];
                            for (int c = 0; c < m; ++c) blk[c] += kl * (long double)wr[c];
                        }
                        for (int c = 0; c < m; ++c) tot[(size_t)r * m4 + c] += blk[c];
                    }
                }
            }
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c < m; ++c) {
                    const long double kw = fast ? (long double)hi[(size_t)r * m4 + c] + (long double)lo[(size_t)r * m4 + c]
                                                : tot[(size_t)r * m4 + c];
                    R[(size_t)c * n + i0 + r] = (double)((long double)B[(size_t)c * n + i0 + r] - kw);
                }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
}

double orc_kernel(int kind, double p0, double p1, double d, int diff) {
    orc::Kern<double> k(kind, p0, p1);
    return diff ? k.diff(d) : k.compute(d);
}

}  // extern "C"
