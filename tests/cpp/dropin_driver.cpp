// dropin_driver.cpp — exercises the drop-in headers the way the reference's callers do
// (tests/test_gaussian.cpp:113-175 and src/gp_node.cpp:898-922, :1074 in the reference): build Data,
// setCovFunction, create<>, the evaluate overloads, update<>.  Results go to a text file that the
// Python tests compare with the oracle.  `--errors` runs the argument checks, which need no GPU.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <gp_regression/gp_regressors.h>

using namespace gp_regression;

template <class Reg, class Cov>
static int run(std::ifstream& in, std::ofstream& out, std::shared_ptr<Cov> kernel) {
    int n, q, k, with_normals;
    in >> n >> q >> k >> with_normals;
    auto data = std::make_shared<Data>();
    for (int i = 0; i < n; ++i) {
        double x, y, z, l, s;
        in >> x >> y >> z >> l >> s;
        data->coord_x.push_back(x); data->coord_y.push_back(y); data->coord_z.push_back(z);
        data->label.push_back(l); data->sigma2.push_back(s);
    }
    auto query = std::make_shared<Data>();
    for (int i = 0; i < q; ++i) {
        double x, y, z;
        in >> x >> y >> z;
        query->coord_x.push_back(x); query->coord_y.push_back(y); query->coord_z.push_back(z);
    }
    auto extra = std::make_shared<Data>();
    for (int i = 0; i < k; ++i) {
        double x, y, z, l, s;
        in >> x >> y >> z >> l >> s;
        extra->coord_x.push_back(x); extra->coord_y.push_back(y); extra->coord_z.push_back(z);
        extra->label.push_back(l); extra->sigma2.push_back(s);
    }
    auto reg = std::make_shared<Reg>();
    reg->setCovFunction(kernel);
    Model::Ptr gp;
    if (with_normals) reg->template create<true>(data, gp); else reg->template create<false>(data, gp);
    out.precision(17);
    out << "R " << gp->R << "\n";
    out << "alpha"; for (int i = 0; i < n; ++i) out << ' ' << gp->alpha(i); out << "\n";
    if (with_normals) { out << "normals"; for (int c = 0; c < 3; ++c) for (int i = 0; i < n; ++i) out << ' ' << gp->N(i, c); out << "\n"; }
    std::vector<double> f1, f2, v2, f3, v3, f4, v4;
    Eigen::MatrixXd N3, N4, Tx, Ty;
    Model::ConstPtr cgp = gp;
    reg->evaluate(cgp, query, f1);
    reg->evaluate(cgp, query, f2, v2);
    reg->evaluate(cgp, query, f3, v3, N3);
    reg->evaluate(cgp, query, f4, v4, N4, Tx, Ty);
    auto dump = [&](const char* name, const std::vector<double>& a) { out << name; for (double x : a) out << ' ' << x; out << "\n"; };
    auto dumpm = [&](const char* name, const Eigen::MatrixXd& m) { out << name; for (int c = 0; c < 3; ++c) for (int i = 0; i < (int)m.rows(); ++i) out << ' ' << m(i, c); out << "\n"; };
    dump("f1", f1); dump("f2", f2); dump("v2", v2); dump("f3", f3); dump("v3", v3); dumpm("N3", N3);
    dump("f4", f4); dump("v4", v4); dumpm("N4", N4); dumpm("Tx", Tx); dumpm("Ty", Ty);
    // single-query calls, the pattern of every real caller (src/gp_node.cpp:1074)
    auto one = std::make_shared<Data>();
    one->coord_x.push_back(query->coord_x[0]); one->coord_y.push_back(query->coord_y[0]); one->coord_z.push_back(query->coord_z[0]);
    std::vector<double> ff, vv;
    reg->evaluate(cgp, one, ff, vv);
    out << "single " << ff[0] << ' ' << vv[0] << "\n";
    {   // extension: the node's per-point sampling loop (src/gp_node.cpp:998-1100) as one batched call
        Data iso; std::vector<double> fi, vi;
        reg->sampleIsoSurface(cgp, 1.01, 0.07, 0.01, iso, fi, vi);
        double vsum = 0; for (double x : vi) vsum += x;
        out << "iso " << iso.coord_x.size() << ' ' << vsum << "\n";
    }
    if (k > 0) {
        reg->template update<false>(extra, gp);
        out << "alpha_updated"; for (int i = 0; i < n + k; ++i) out << ' ' << gp->alpha(i); out << "\n";
        out << "R_updated " << gp->R << "\n";
        std::vector<double> fu;
        reg->evaluate(cgp, query, fu);
        dump("f_updated", fu);
    }
    return 0;
}

static std::string message_of(int which) {
    ThinPlateRegressor reg;
    Model::Ptr m;
    std::vector<double> f;
    try {
        if (which == 0) { Data::Ptr d; reg.create<false>(d, m); }
        else if (which == 1) { auto d = std::make_shared<Data>(); reg.create<false>(d, m); }
        else if (which == 2) { auto d = std::make_shared<Data>(); d->coord_x.push_back(0); d->coord_y.push_back(0); d->coord_z.push_back(0); Model::ConstPtr c; reg.evaluate(c, d, f); }
        else if (which == 3) { auto d = std::make_shared<Data>(); d->coord_x.push_back(0); d->coord_y.push_back(0); d->coord_z.push_back(0); Model::Ptr none; reg.update<false>(d, none); }
        else if (which == 4) {
            // labelled query against a (device-less) model: the label check comes before any GPU work
            auto d = std::make_shared<Data>(); d->coord_x.push_back(0); d->coord_y.push_back(0); d->coord_z.push_back(0); d->label.push_back(1);
            Model::ConstPtr c = std::make_shared<Model>(); reg.evaluate(c, d, f);
        }
    } catch (const std::exception& e) { return e.what(); }
    return "";
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "--errors")) {
        for (int w = 0; w < 5; ++w) std::cout << w << ": " << message_of(w) << "\n";
        Eigen::Vector3d N, Tx, Ty;
        computeTangentBasis(Eigen::Vector3d(0, 0, 2), N, Tx, Ty);
        std::cout << "basis " << N(0) << ' ' << N(1) << ' ' << N(2) << ' ' << Tx(0) << ' ' << Tx(1) << ' ' << Tx(2) << ' '
                  << Ty(0) << ' ' << Ty(1) << ' ' << Ty(2) << "\n";
        computeTangentBasis(Eigen::Vector3d(3, 0, 0), N, Tx, Ty);
        std::cout << "basis " << N(0) << ' ' << N(1) << ' ' << N(2) << ' ' << Tx(0) << ' ' << Tx(1) << ' ' << Tx(2) << ' '
                  << Ty(0) << ' ' << Ty(1) << ' ' << Ty(2) << "\n";
        ThinPlate tp(2.0); Gaussian ga; Laplace la(1.5, 0.5);
        double one = 1.0;
        std::cout << "kern " << tp.compute(1.0) << ' ' << tp.computediff(1.0) << ' ' << tp.compute(0.0) << ' ' << tp.compute(2.0) << ' '
                  << ga.compute(one) << ' ' << ga.computediff(one) << ' ' << la.compute(one) << ' ' << la.computediff(one) << "\n";
        return 0;
    }
    if (argc >= 3 && !std::strcmp(argv[1], "--pcd")) {          // host only: the PCD reader behind loadPCD
        try {
            Data::Ptr d = loadPCD(argv[2]);
            double sx = 0, sy = 0, sz = 0;
            for (size_t i = 0; i < d->coord_x.size(); ++i) { sx += d->coord_x[i]; sy += d->coord_y[i]; sz += d->coord_z[i]; }
            std::cout.precision(17);
            std::cout << "pcd " << d->coord_x.size() << ' ' << sx << ' ' << sy << ' ' << sz << ' ' << d->label.size() << "\n";
        } catch (const GPRegressionException& e) { std::cout << "exception " << e.what() << "\n"; }
        return 0;
    }
    if (argc < 3) { std::cerr << "usage: dropin_driver <in.txt> <out.txt> | --errors | --pcd <file.pcd>\n"; return 2; }
    std::ifstream in(argv[1]);
    std::ofstream out(argv[2]);
    int kind; double p0, p1;
    in >> kind >> p0 >> p1;
    try {
        if (kind == 0) return run<ThinPlateRegressor, ThinPlate>(in, out, std::make_shared<ThinPlate>(p0));
        if (kind == 1) return run<GaussianRegressor, Gaussian>(in, out, std::make_shared<Gaussian>(p0, p1));
        return run<LaplaceRegressor, Laplace>(in, out, std::make_shared<Laplace>(p0, p1));
    } catch (const GPRegressionException& e) {
        out << "exception " << e.status() << ' ' << e.pivot() << ' ' << e.what() << "\n";
        return 0;
    }
}
