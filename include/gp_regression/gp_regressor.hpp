// Drop-in for /root/reference/include/gp_regression/gp_regressor.hpp (namespace gp_regression):
// Data, Model, GPRegressor<CovType> with create<bool> / four evaluate overloads / update<bool> /
// setCovFunction, and the free function computeTangentBasis — same names, signatures, ownership and
// exception messages.  All arithmetic runs in libgpr_b200.so (hand-written sm_100a CUDA behind
// include/gpr_c_api.h); this header only validates, marshals the SoA vectors and copies results back.
// There is no CPU fallback: without the library / a B200 the calls throw GPRegressionException.
//
// Differences a caller can observe (documented in INTEGRATION.md):
//   * Model::Kpp, Kppdiff, Kppdiffdiff stay empty and there is no Model::cholesker — no caller reads
//     them; the factor lives on the device behind Model::device.
//   * an indefinite covariance (the node's ThinPlate(2.0) + external sphere setting) is factorised
//     as a block L D L^T with one dense trailing pivot block of up to 256 offending points, so create() succeeds
//     and agrees with Eigen's pivoted LDLT to cond*eps; only beyond that limit (or for a singular matrix) does
//     create() throw GPRegressionException("covariance matrix is not positive definite …").
//   * outputs N are zero-initialised before accumulation (the reference adds into uninitialised
//     storage, gp_regressor.hpp:241/:247).
//   * computeTangentBasis is inline (the reference defines it non-inline in a header, :29).
#pragma once

#include <cstdlib>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(GPR_FORCE_MINI_EIGEN)
#include <Eigen/Core>
#define GPR_HAVE_EIGEN 1
#endif
#endif
#ifndef GPR_HAVE_EIGEN
#include "mini_eigen.h"
#endif

#include "../gpr_c_api.h"
#include "cov_functions.h"
#include "gp_regression_exception.h"

namespace gp_regression {

// gp_regressor.hpp:29-44.  N = grad/|grad|; Tx = e - N (N.e) normalised with e = X unless N is within
// 1e-3 of X (Eigen isApprox), then e = Y; Ty = N x Tx normalised.
inline void computeTangentBasis(const Eigen::Vector3d& grad, Eigen::Vector3d& N, Eigen::Vector3d& Tx,
                                Eigen::Vector3d& Ty) {
    N = grad.normalized();
    const double dx = N(0) - 1.0;
    const double diff2 = dx * dx + N(1) * N(1) + N(2) * N(2);
    const double nn = N(0) * N(0) + N(1) * N(1) + N(2) * N(2);
    const bool near_x = diff2 <= 1e-6 * (nn < 1.0 ? nn : 1.0);
    Eigen::Vector3d e = near_x ? Eigen::Vector3d(0, 1, 0) : Eigen::Vector3d(1, 0, 0);
    const double ne = N.dot(e);
    Tx = Eigen::Vector3d(e(0) - N(0) * ne, e(1) - N(1) * ne, e(2) - N(2) * ne);
    Tx.normalize();
    Ty = N.cross(Tx);
    Ty.normalize();
}

// gp_regressor.hpp:49-66
struct Data {
    std::vector<double> coord_x, coord_y, coord_z, label, sigma2;
    typedef std::shared_ptr<Data> Ptr;
    typedef std::shared_ptr<const Data> ConstPtr;
    void clear() { coord_x.clear(); coord_y.clear(); coord_z.clear(); label.clear(); sigma2.clear(); }
};

// gp_regressor.hpp:71-87
struct Model {
    double R = 0.0;             // largest pairwise training distance
    Eigen::MatrixXd P;          // n x 3 points
    Eigen::VectorXd Y, S2;      // labels, noise
    Eigen::MatrixXd N;          // normals at the training points (create<true>)
    Eigen::MatrixXd Tx, Ty;     // never filled by the reference either
    Eigen::MatrixXd Kpp;        // left empty: K is assembled and factorised on the device
    Eigen::VectorXd alpha;      // weights
    Eigen::MatrixXd Kppdiff, Kppdiffdiff;   // left empty
    std::shared_ptr<void> device;           // gpr_model*, freed with the last Model::Ptr
    typedef std::shared_ptr<Model> Ptr;
    typedef std::shared_ptr<const Model> ConstPtr;
};

namespace detail {

inline void raise(int rc) {
    throw GPRegressionException(gpr_last_error(), rc, rc == GPR_ERR_NOT_SPD ? gpr_last_pivot() : 0);
}

// One process-wide context, created on first use.  GPR_DEVICES="0,1,2,3" selects the GPUs that share
// the query load (default: device 0).
inline gpr_ctx* context() {
    static gpr_ctx* ctx = nullptr;
    static std::once_flag once;
    static int rc = 0;
    std::call_once(once, [] {
        std::vector<int> devs;
        if (const char* s = std::getenv("GPR_DEVICES")) {
            std::stringstream ss(s);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) devs.push_back(std::atoi(tok.c_str()));
        }
        rc = gpr_ctx_create(devs.empty() ? nullptr : devs.data(), (int)devs.size(), &ctx);
    });
    if (!ctx) throw GPRegressionException(std::string("cannot create GPU context: ") + gpr_last_error(), rc ? rc : GPR_ERR_CUDA);
    return ctx;
}

inline gpr_model* handle(const Model& m) { return static_cast<gpr_model*>(m.device.get()); }

inline void vec_to(Eigen::VectorXd& dst, const std::vector<double>& src) {
    dst.resize((Eigen::Index)src.size());
    for (size_t i = 0; i < src.size(); ++i) dst((Eigen::Index)i) = src[i];
}

}  // namespace detail

// Extension: the x, y, z fields of a PCD file (ascii / binary / binary_compressed) as a Data with empty labels — what the
// reference's node obtains from PCL (pcl::io::loadPCDFile, src/gp_node.cpp:557); no PCL needed (gpr_pcd_read_xyz).
inline Data::Ptr loadPCD(const std::string& path) {
    double *x = nullptr, *y = nullptr, *z = nullptr;
    size_t n = 0;
    const int rc = gpr_pcd_read_xyz(path.c_str(), &x, &y, &z, &n);
    if (rc != GPR_OK) detail::raise(rc);
    Data::Ptr d = std::make_shared<Data>();
    d->coord_x.assign(x, x + n); d->coord_y.assign(y, y + n); d->coord_z.assign(z, z + n);
    gpr_free(x); gpr_free(y); gpr_free(z);
    return d;
}

template <typename CovType>
class GPRegressor {
public:
    std::shared_ptr<CovType> kernel_;       // gp_regressor.hpp:97

    virtual ~GPRegressor() {}
    GPRegressor() : kernel_(std::make_shared<CovType>()) {}                             // :497-500
    void setCovFunction(const std::shared_ptr<CovType>& kernel) { kernel_ = kernel; }   // :488-491

    // :110-182.  Replaces gp with a fresh Model (":117 we dont care what there was there").
    template <bool withNormals>
    void create(Data::ConstPtr data, Model::Ptr& gp) {
        assertData(data);
        const size_t n = data->coord_x.size();
        if (data->coord_y.size() != n || data->coord_z.size() != n || data->label.size() != n ||
            (!data->sigma2.empty() && data->sigma2.size() < n))
            throw GPRegressionException("Inconsistent input data sizes");
        gp = std::make_shared<Model>();
        gpr_model* h = nullptr;
        const int rc = gpr_fit(detail::context(), data->coord_x.data(), data->coord_y.data(), data->coord_z.data(),
                               data->label.data(), data->sigma2.empty() ? nullptr : data->sigma2.data(), n,
                               kernel_->descriptor(), withNormals ? 1 : 0, &h);
        if (rc != GPR_OK) detail::raise(rc);
        gp->device = std::shared_ptr<void>(h, [](void* p) { gpr_model_destroy(static_cast<gpr_model*>(p)); });
        refreshHostMirror(*data, *gp, withNormals, 0);
    }

    // :194-212
    void evaluate(Model::ConstPtr gp, Data::ConstPtr query, std::vector<double>& f, std::vector<double>& v,
                  Eigen::MatrixXd& N, Eigen::MatrixXd& Tx, Eigen::MatrixXd& Ty) {
        run(gp, query, f, &v, &N, &Tx, &Ty);
    }
    // :222-273  (N is the un-normalised gradient, :250)
    void evaluate(Model::ConstPtr gp, Data::ConstPtr query, std::vector<double>& f, std::vector<double>& v,
                  Eigen::MatrixXd& N) {
        run(gp, query, f, &v, &N, nullptr, nullptr);
    }
    // :282-324
    void evaluate(Model::ConstPtr gp, Data::ConstPtr query, std::vector<double>& f, std::vector<double>& v) {
        run(gp, query, f, &v, nullptr, nullptr, nullptr);
    }
    // :332-357
    void evaluate(Model::ConstPtr gp, Data::ConstPtr query, std::vector<double>& f) {
        run(gp, query, f, nullptr, nullptr, nullptr, nullptr);
    }

    // :367-479.  Appends new_data and re-solves; R and the normals are not refreshed (:454-455, :462-477).
    template <bool withNormals>
    void update(Data::ConstPtr new_data, Model::Ptr gp) {
        assertData(new_data);
        if (!gp) throw GPRegressionException("Empty model pointer");
        if (!gp->device) throw GPRegressionException("Model was not created by this regressor");
        const size_t k = new_data->label.size();                                       // :383
        if (new_data->coord_x.size() != k || new_data->coord_y.size() != k || new_data->coord_z.size() != k ||
            (!new_data->sigma2.empty() && new_data->sigma2.size() < k))
            throw GPRegressionException("Inconsistent input data sizes");
        const size_t p = (size_t)gp->Y.size();
        const int rc = gpr_append(detail::context(), detail::handle(*gp), new_data->coord_x.data(),
                                  new_data->coord_y.data(), new_data->coord_z.data(), new_data->label.data(),
                                  new_data->sigma2.empty() ? nullptr : new_data->sigma2.data(), k);
        if (rc != GPR_OK) detail::raise(rc);
        refreshHostMirror(*new_data, *gp, false, p);
    }

    // Extension (no counterpart in the reference's regressor): the node's fakeDeterministicSampling /
    // samplePoint loop (src/gp_node.cpp:998-1100 — one thread and one evaluate(q = 1) per lattice point of
    // [-scale, scale]^3 with spacing `pass`, keep |f| <= tol, variance as intensity) as ONE batched call.
    // points / f / v receive the kept lattice points in lattice order.
    void sampleIsoSurface(Model::ConstPtr gp, double scale, double pass, double tol, Data& points,
                          std::vector<double>& f, std::vector<double>& v) {
        if (!gp) throw GPRegressionException("Empty Model pointer");
        if (!gp->device) throw GPRegressionException("Model was not created by this regressor");
        // One pass over the lattice with a capacity guess (a thin shell: ~1/16 of the lattice); the lattice is only
        // evaluated again when the guess was too small.
        size_t na = 0;
        for (double a = -scale; a <= scale; a += pass) ++na;
        size_t cap = na * na * na / 16 + 1024, count = 0;
        points.clear();
        for (int attempt = 0; attempt < 2; ++attempt) {
            points.coord_x.assign(cap, 0.0); points.coord_y.assign(cap, 0.0); points.coord_z.assign(cap, 0.0);
            f.assign(cap, 0.0); v.assign(cap, 0.0);
            const int rc = gpr_sample_isosurface(detail::context(), detail::handle(*gp), -scale, scale, pass, tol, cap,
                                                 points.coord_x.data(), points.coord_y.data(), points.coord_z.data(), f.data(),
                                                 v.data(), &count);
            if (rc != GPR_OK) detail::raise(rc);
            if (count <= cap) break;
            cap = count;
        }
        points.coord_x.resize(count); points.coord_y.resize(count); points.coord_z.resize(count);
        f.resize(count); v.resize(count);
    }

    // Extension: AtlasVariance::sampleOnChart (include/atlas/atlas_variance.hpp:147-219) for any number of charts in one
    // call.  centers / normals / tx / ty: c x 3 (Chart::getCenter / getNormal / getTanBasisOne / getTanBasisTwo), radii: c,
    // counts: samples per chart (the reference: ceil(|disc_samples_factor| * R)).  On return samples[k] (counts[k] x 3) and
    // vars_ids[k] are what the reference leaves in Chart::samples and Chart::vars_ids (sorted by decreasing variance).
    // The uniform variates come from mt19937_64(seed) with the reference's distributions (random_generation.hpp:16-27).
    void sampleOnCharts(Model::ConstPtr gp, const Eigen::MatrixXd& centers, const Eigen::MatrixXd& normals,
                        const Eigen::MatrixXd& tx, const Eigen::MatrixXd& ty, const std::vector<double>& radii,
                        const std::vector<size_t>& counts, unsigned long long seed, std::vector<Eigen::MatrixXd>& samples,
                        std::vector<std::vector<std::pair<double, std::size_t> > >& vars_ids) {
        if (!gp) throw GPRegressionException("Empty Model pointer");
        if (!gp->device) throw GPRegressionException("Model was not created by this regressor");
        const size_t c = (size_t)centers.rows();
        if (c == 0 || centers.cols() != 3 || (size_t)normals.rows() != c || (size_t)tx.rows() != c || (size_t)ty.rows() != c ||
            radii.size() != c || counts.size() != c)
            throw GPRegressionException("Inconsistent input data sizes");
        std::vector<double> frames(13 * c);
        size_t total = 0;
        for (size_t k = 0; k < c; ++k) {
            for (int a = 0; a < 3; ++a) {
                frames[13 * k + a] = centers((Eigen::Index)k, a); frames[13 * k + 3 + a] = normals((Eigen::Index)k, a);
                frames[13 * k + 6 + a] = tx((Eigen::Index)k, a); frames[13 * k + 9 + a] = ty((Eigen::Index)k, a);
            }
            frames[13 * k + 12] = radii[k];
            total += counts[k];
        }
        std::vector<double> sx(total), sy(total), sz(total), f(total), v(total);
        std::vector<size_t> order(total);
        const int rc = gpr_sample_chart(detail::context(), detail::handle(*gp), frames.data(), counts.data(), c, nullptr, nullptr,
                                        seed, sx.data(), sy.data(), sz.data(), f.data(), v.data(), order.data());
        if (rc != GPR_OK) detail::raise(rc);
        samples.assign(c, Eigen::MatrixXd());
        vars_ids.assign(c, std::vector<std::pair<double, std::size_t> >());
        size_t o = 0;
        for (size_t k = 0; k < c; ++k) {
            samples[k].resize((Eigen::Index)counts[k], 3);
            for (size_t i = 0; i < counts[k]; ++i) {
                samples[k]((Eigen::Index)i, 0) = sx[o + i]; samples[k]((Eigen::Index)i, 1) = sy[o + i]; samples[k]((Eigen::Index)i, 2) = sz[o + i];
                vars_ids[k].push_back(std::make_pair(v[o + order[o + i]], order[o + i]));
            }
            o += counts[k];
        }
    }

    // Extension: AtlasBase::project (include/atlas/atlas.hpp:201-276, same defaults) for any number of points in one
    // call.  in / normals / out: k x 3 (rows = points; normals = initial un-normalised gradients).  Returns, per
    // point, the iterations used when a tolerance was met or -max_iter when the budget ran out.
    std::vector<int> projectOnSurface(Model::ConstPtr gp, const Eigen::MatrixXd& in, const Eigen::MatrixXd& normals,
                                      Eigen::MatrixXd& out, double f_tol = 1e-2, double improve_tol = 1e-7,
                                      unsigned int max_iter = 500, double step_mul = 0.001) {
        if (!gp) throw GPRegressionException("Empty Model pointer");
        if (!gp->device) throw GPRegressionException("Model was not created by this regressor");
        const size_t k = (size_t)in.rows();
        if (k == 0 || in.cols() != 3 || normals.rows() != in.rows() || normals.cols() != 3)
            throw GPRegressionException("All input data is empty!");
        std::vector<double> buf(9 * k);
        for (size_t i = 0; i < k; ++i)
            for (int c = 0; c < 3; ++c) { buf[c * k + i] = in((Eigen::Index)i, c); buf[(3 + c) * k + i] = normals((Eigen::Index)i, c); }
        std::vector<int> status(k);
        const int rc = gpr_project(detail::context(), detail::handle(*gp), &buf[0], &buf[k], &buf[2 * k], &buf[3 * k], &buf[4 * k],
                                   &buf[5 * k], k, f_tol, improve_tol, max_iter, step_mul, &buf[6 * k], &buf[7 * k], &buf[8 * k],
                                   status.data());
        if (rc != GPR_OK) detail::raise(rc);
        out.resize((Eigen::Index)k, 3);
        for (size_t i = 0; i < k; ++i)
            for (int c = 0; c < 3; ++c) out((Eigen::Index)i, c) = buf[(6 + c) * k + i];
        return status;
    }

private:
    // :563-572
    void assertData(Data::ConstPtr data) const {
        if (!data) throw GPRegressionException("Empty data pointer");
        if (data->coord_x.empty() && data->coord_y.empty() && data->coord_z.empty() && data->label.empty())
            throw GPRegressionException("All input data is empty!");
    }

    // Host mirrors of the fields the reference's Model exposes: P, Y, S2 (appended from `added`
    // starting at row `first`), alpha, R, N.
    void refreshHostMirror(const Data& added, Model& gp, bool normals, size_t first) const {
        gpr_model* h = detail::handle(gp);
        const size_t n = gpr_model_size(h);
        Eigen::MatrixXd P((Eigen::Index)n, 3);
        Eigen::VectorXd Y, S2;
        Y.resize((Eigen::Index)n);
        const bool s2 = !added.sigma2.empty() || gp.S2.size() > 0;
        if (s2) S2.resize((Eigen::Index)n);
        for (size_t i = 0; i < first; ++i) {
            for (int c = 0; c < 3; ++c) P((Eigen::Index)i, c) = gp.P((Eigen::Index)i, c);
            Y((Eigen::Index)i) = gp.Y((Eigen::Index)i);
            if (s2) S2((Eigen::Index)i) = gp.S2.size() > (Eigen::Index)i ? gp.S2((Eigen::Index)i) : 0.0;
        }
        for (size_t i = first; i < n; ++i) {
            const size_t s = i - first;
            P((Eigen::Index)i, 0) = added.coord_x[s]; P((Eigen::Index)i, 1) = added.coord_y[s]; P((Eigen::Index)i, 2) = added.coord_z[s];
            Y((Eigen::Index)i) = added.label[s];
            if (s2) S2((Eigen::Index)i) = s < added.sigma2.size() ? added.sigma2[s] : 0.0;
        }
        gp.P = P; gp.Y = Y; gp.S2 = S2;
        std::vector<double> alpha(n), nrm(normals ? 3 * n : 0);
        double R = 0.0;
        const int rc = gpr_model_get(h, alpha.data(), &R, normals ? nrm.data() : nullptr);
        if (rc != GPR_OK) detail::raise(rc);
        detail::vec_to(gp.alpha, alpha);
        gp.R = R;
        if (normals) {
            gp.N.resize((Eigen::Index)n, 3);
            for (size_t i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) gp.N((Eigen::Index)i, c) = nrm[(size_t)c * n + i];
        }
    }

    void run(const Model::ConstPtr& gp, const Data::ConstPtr& query, std::vector<double>& f, std::vector<double>* v,
             Eigen::MatrixXd* N, Eigen::MatrixXd* Tx, Eigen::MatrixXd* Ty) {
        if (!gp) throw GPRegressionException("Empty Model pointer");                    // :197, :224, :284, :334
        assertData(query);                                                             // :228, :288, :338
        if (!query->label.empty()) throw GPRegressionException("Query is already labeled!");   // :230, :290, :340
        if (!gp->device) throw GPRegressionException("Model was not created by this regressor");
        const size_t q = query->coord_x.size();
        if (query->coord_y.size() != q || query->coord_z.size() != q)
            throw GPRegressionException("Inconsistent query sizes");
        f.assign(q, 0.0);
        if (v) v->assign(q, 0.0);
        if (N) N->resize((Eigen::Index)q, 3);
        if (Tx) Tx->resize((Eigen::Index)q, 3);
        if (Ty) Ty->resize((Eigen::Index)q, 3);
        const int rc = gpr_predict(detail::context(), detail::handle(*gp), query->coord_x.data(), query->coord_y.data(),
                                   query->coord_z.data(), q, f.data(), v ? v->data() : nullptr, N ? N->data() : nullptr,
                                   Tx ? Tx->data() : nullptr, Ty ? Ty->data() : nullptr);
        if (rc != GPR_OK) detail::raise(rc);
    }
};

}  // namespace gp_regression
