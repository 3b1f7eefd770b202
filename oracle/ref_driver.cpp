// ref_driver.cpp — C interface around the reference's OWN, unmodified header
// /root/reference/include/gp_regression/gp_regressors.h, compiled where it lies against the
// Eigen API shim in oracle/eigen_shim (Eigen itself is absent from this image).
//
// TEST INFRASTRUCTURE ONLY.  Output goes to oracle/_ref/libgpr_ref.so (git-ignored, travels
// to the GPU box).  It is used (a) to pin oracle/gpr_oracle.cpp and (b) by
// oracle/make_golden.py to generate the fixtures committed under tests/golden/.
// No reference source is copied into this repository: this file only *includes* it.
#include <memory>
#include <string>
#include <vector>
#include <cstring>
#include <thread>
#include <algorithm>

#include <gp_regression/gp_regressors.h>   // -> /root/reference/include (reference, unmodified)

using namespace gp_regression;

namespace {
struct Handle {
    int kind;
    std::shared_ptr<ThinPlateRegressor> tp;
    std::shared_ptr<GaussianRegressor> ga;
    std::shared_ptr<LaplaceRegressor> la;
    Model::Ptr model;
    std::string err;
};

Data::Ptr make_data(const double* x, const double* y, const double* z, const double* label,
                    const double* sigma2, int n) {
    auto d = std::make_shared<Data>();
    d->coord_x.assign(x, x + n); d->coord_y.assign(y, y + n); d->coord_z.assign(z, z + n);
    if (label) d->label.assign(label, label + n);
    if (sigma2) d->sigma2.assign(sigma2, sigma2 + n);
    return d;
}

template <class F> int guarded(Handle* h, F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { h->err = e.what(); return 1; }
}
}  // namespace

extern "C" {

void* ref_create(int kind, double p0, double p1) {
    Handle* h = new Handle();
    h->kind = kind;
    if (kind == 0) { h->tp = std::make_shared<ThinPlateRegressor>(); h->tp->setCovFunction(std::make_shared<ThinPlate>(p0)); }
    else if (kind == 1) { h->ga = std::make_shared<GaussianRegressor>(); h->ga->setCovFunction(std::make_shared<Gaussian>(p0, p1)); }
    else { h->la = std::make_shared<LaplaceRegressor>(); h->la->setCovFunction(std::make_shared<Laplace>(p0, p1)); }
    return h;
}
void ref_free(void* p) { delete (Handle*)p; }
const char* ref_error(void* p) { return ((Handle*)p)->err.c_str(); }

int ref_fit(void* p, const double* x, const double* y, const double* z, const double* label,
            const double* sigma2_or_null, int n, int with_normals) {
    Handle* h = (Handle*)p;
    return guarded(h, [&] {
        Data::Ptr d = make_data(x, y, z, label, sigma2_or_null, n);
        if (h->kind == 0) { if (with_normals) h->tp->create<true>(d, h->model); else h->tp->create<false>(d, h->model); }
        else if (h->kind == 1) { if (with_normals) h->ga->create<true>(d, h->model); else h->ga->create<false>(d, h->model); }
        else { if (with_normals) h->la->create<true>(d, h->model); else h->la->create<false>(d, h->model); }
    });
}

int ref_n(void* p) { return (int)((Handle*)p)->model->alpha.size(); }
double ref_R(void* p) { return ((Handle*)p)->model->R; }

// alpha[n]; normals n x 3 column-major (only after with_normals); K n x n column-major.
void ref_get(void* p, double* alpha, double* normals, double* K) {
    Handle* h = (Handle*)p;
    const Model& m = *h->model;
    if (alpha) std::memcpy(alpha, m.alpha.data(), sizeof(double) * m.alpha.size());
    if (normals && m.N.size()) std::memcpy(normals, m.N.data(), sizeof(double) * m.N.size());
    if (K) std::memcpy(K, m.Kpp.data(), sizeof(double) * m.Kpp.size());
}

// mode 1: f; 2: f,v; 3: f,v,N(grad); 4: f,v,N,Tx,Ty.   Matrices q x 3 column-major.
int ref_evaluate(void* p, const double* qx, const double* qy, const double* qz, int q, int mode,
                 double* f, double* v, double* N, double* Tx, double* Ty) {
    Handle* h = (Handle*)p;
    return guarded(h, [&] {
        Data::Ptr d = make_data(qx, qy, qz, nullptr, nullptr, q);
        std::vector<double> ff, vv;
        Eigen::MatrixXd NN, TX, TY;
        Model::ConstPtr gp = h->model;
        auto run = [&](auto& reg) {
            if (mode == 1) reg.evaluate(gp, d, ff);
            else if (mode == 2) reg.evaluate(gp, d, ff, vv);
            else if (mode == 3) reg.evaluate(gp, d, ff, vv, NN);
            else reg.evaluate(gp, d, ff, vv, NN, TX, TY);
        };
        if (h->kind == 0) run(*h->tp); else if (h->kind == 1) run(*h->ga); else run(*h->la);
        std::memcpy(f, ff.data(), sizeof(double) * q);
        if (mode >= 2) std::memcpy(v, vv.data(), sizeof(double) * q);
        if (mode >= 3) std::memcpy(N, NN.data(), sizeof(double) * 3 * q);
        if (mode >= 4) { std::memcpy(Tx, TX.data(), sizeof(double) * 3 * q); std::memcpy(Ty, TY.data(), sizeof(double) * 3 * q); }
    });
}

int ref_update(void* p, const double* x, const double* y, const double* z, const double* label,
               const double* sigma2_or_null, int k) {
    Handle* h = (Handle*)p;
    return guarded(h, [&] {
        Data::Ptr d = make_data(x, y, z, label, sigma2_or_null, k);
        if (h->kind == 0) h->tp->update<false>(d, h->model);
        else if (h->kind == 1) h->ga->update<false>(d, h->model);
        else h->la->update<false>(d, h->model);
    });
}

// Timing support for bench.py --impl reference (NOT used by any parity test): builds the reference's Model
// from an externally computed Cholesky factor (the reference's own LDLT::compute is unblocked and
// single-threaded: tens of minutes at n = 16384), so that the reference's UNMODIFIED evaluate() can be
// timed at the bench size.  Fields filled: P, Y, S2, alpha, R, cholesker (gp_regressor.hpp:71-87).
int ref_adopt(void* p, const double* x, const double* y, const double* z, const double* label,
              const double* sigma2_or_null, int n, const double* alpha, const double* Lc, double R) {
    Handle* h = (Handle*)p;
    return guarded(h, [&] {
        auto m = std::make_shared<Model>();
        m->P.resize(n, 3);
        m->Y.resize(n);
        m->alpha.resize(n);
        if (sigma2_or_null) m->S2.resize(n);
        for (int i = 0; i < n; ++i) {
            m->P(i, 0) = x[i]; m->P(i, 1) = y[i]; m->P(i, 2) = z[i];
            m->Y(i) = label[i]; m->alpha(i) = alpha[i];
            if (sigma2_or_null) m->S2(i) = sigma2_or_null[i];
        }
        m->R = R;
        m->cholesker.adoptCholesky(Lc, n);
        h->model = m;
    });
}

// The reference's evaluate() called concurrently from `threads` std::threads on one shared const Model, each
// on its own contiguous chunk of the queries with `per_call` queries per call — the way the node drives it
// (one thread per grid point, q = 1 per call, src/gp_node.cpp:1027-1038, :1074).  mode 1: f; 2: f, v.
int ref_evaluate_mt(void* p, const double* qx, const double* qy, const double* qz, int q, int mode, int threads,
                    int per_call, double* f, double* v) {
    Handle* h = (Handle*)p;
    if (threads < 1) threads = 1;
    if (per_call < 1) per_call = 1;
    std::vector<std::string> errs((size_t)threads);
    std::vector<std::thread> pool;
    Model::ConstPtr gp = h->model;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            const int a = (int)((long long)q * t / threads), b = (int)((long long)q * (t + 1) / threads);
            try {
                for (int s = a; s < b; s += per_call) {
                    const int c = std::min(per_call, b - s);
                    Data::Ptr d = make_data(qx + s, qy + s, qz + s, nullptr, nullptr, c);
                    std::vector<double> ff, vv;
                    auto run = [&](auto& reg) { if (mode == 1) reg.evaluate(gp, d, ff); else reg.evaluate(gp, d, ff, vv); };
                    if (h->kind == 0) run(*h->tp); else if (h->kind == 1) run(*h->ga); else run(*h->la);
                    std::memcpy(f + s, ff.data(), sizeof(double) * c);
                    if (mode >= 2) std::memcpy(v + s, vv.data(), sizeof(double) * c);
                }
            } catch (const std::exception& e) { errs[(size_t)t] = e.what(); }
        });
    }
    for (auto& th : pool) th.join();
    for (auto& e : errs) if (!e.empty()) { h->err = e; return 1; }
    return 0;
}

// Exercise the reference's argument checks (gp_regressor.hpp:197,224,230,284,290,334,340,373,563-572).
// which: 0 null data to create, 1 all-empty data to create, 2 labelled query, 3 null model to evaluate.
const char* ref_error_message(int which) {
    static std::string msg;
    msg.clear();
    ThinPlateRegressor reg;
    Model::Ptr m;
    std::vector<double> f;
    try {
        if (which == 0) { Data::Ptr d; reg.create<false>(d, m); }
        else if (which == 1) { auto d = std::make_shared<Data>(); reg.create<false>(d, m); }
        else if (which == 2) {
            double x = 0, y = 0, z = 0, l = 1, s = 0.1;
            auto d = make_data(&x, &y, &z, &l, &s, 1); reg.create<false>(d, m);
            reg.evaluate(m, d, f);
        } else { double x = 0; auto d = make_data(&x, &x, &x, nullptr, nullptr, 1); Model::ConstPtr c; reg.evaluate(c, d, f); }
    } catch (const std::exception& e) { msg = e.what(); }
    return msg.c_str();
}

}  // extern "C"
