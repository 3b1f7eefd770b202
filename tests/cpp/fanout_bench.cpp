// fanout_bench.cpp — the UNCHANGED caller: the node's grid sampling loop, one std::thread and one
// evaluate(gp, q, f, v) with a single query per lattice point, all threads of a slab sharing one regressor and
// one const Model (reference: src/gp_node.cpp:1025-1038 fakeDeterministicSampling, :1067-1100 samplePoint).
//
// This one source is compiled twice, against two header sets with the same API:
//   * include/gp_regression (this repository's drop-in headers over libgpr_b200.so)      -> the GPU arm
//   * /root/reference/include/gp_regression (the reference's own header, oracle/eigen_shim) -> the CPU arm
//     (built by oracle/Makefile into oracle/_ref/fanout_ref, which travels to the GPU box)
// so the two arms run literally the same caller code on the same box.
//
//   fanout_bench <input.txt> <out.bin> [scale=1.01] [pass=0.07]
// input: R, n, then n rows "x y z label sigma2".  out.bin: lattice count (int64), then f[count], v[count] in
// lattice order.  stdout: one JSON line with the timings.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <gp_regression/gp_regressors.h>

using namespace gp_regression;

struct Sampler {
    ThinPlateRegressor::Ptr reg_;
    Model::Ptr obj_gp;
    std::mutex mtx_samp;
    std::vector<double> kept_x, kept_y, kept_z, kept_v;      // real_explicit_ptr of the node
    double *f_all, *v_all;

    // src/gp_node.cpp:1067-1100 without the ROS marker bookkeeping
    void samplePoint(const double x, const double y, const double z, size_t slot) {
        Data::Ptr qq = std::make_shared<Data>();
        qq->coord_x.push_back(x);
        qq->coord_y.push_back(y);
        qq->coord_z.push_back(z);
        std::vector<double> ff, vv;
        reg_->evaluate(obj_gp, qq, ff, vv);
        f_all[slot] = ff.at(0);
        v_all[slot] = vv.at(0);
        if (std::abs(ff.at(0)) <= 0.01) {
            std::lock_guard<std::mutex> lk(mtx_samp);
            kept_x.push_back(x); kept_y.push_back(y); kept_z.push_back(z); kept_v.push_back(vv[0]);
        }
    }
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: fanout_bench <input.txt> <out.bin> [scale] [pass]\n"); return 2; }
    const double scale = argc > 3 ? std::atof(argv[3]) : 1.01, pass = argc > 4 ? std::atof(argv[4]) : 0.07;
    std::ifstream in(argv[1]);
    double R; int n;
    in >> R >> n;
    Data::Ptr data = std::make_shared<Data>();
    for (int i = 0; i < n; ++i) {
        double x, y, z, l, s;
        in >> x >> y >> z >> l >> s;
        data->coord_x.push_back(x); data->coord_y.push_back(y); data->coord_z.push_back(z);
        data->label.push_back(l); data->sigma2.push_back(s);
    }
    if (!in) { std::fprintf(stderr, "bad input file\n"); return 2; }
    Sampler S;
    S.reg_ = std::make_shared<ThinPlateRegressor>();
    S.reg_->setCovFunction(std::make_shared<ThinPlate>(R));          // src/gp_node.cpp:917-920
    const double t_fit0 = now_s();
    S.reg_->create<false>(data, S.obj_gp);                            // :922
    const double fit_s = now_s() - t_fit0;

    std::vector<double> axis;
    for (double a = -scale; a <= scale; a += pass) axis.push_back(a);
    const size_t na = axis.size(), total = na * na * na;
    std::vector<double> f_all(total), v_all(total);
    S.f_all = f_all.data(); S.v_all = v_all.data();

    // the cost of the caller's own thread management, identical in both arms: spawn + join of na*na empty threads
    double spawn_s = 0.0;
    {
        const double t0 = now_s();
        std::vector<std::thread> threads;
        for (size_t i = 0; i < na * na; ++i) threads.emplace_back([] {});
        for (auto& t : threads) t.join();
        spawn_s = now_s() - t0;
    }
    std::vector<double> slab_ms;
    const double t0 = now_s();
    size_t slot = 0;
    for (double x = -scale; x <= scale; x += pass) {                  // :1025-1038
        const double ts = now_s();
        std::vector<std::thread> threads;
        for (double y = -scale; y <= scale; y += pass)
            for (double z = -scale; z <= scale; z += pass)
                threads.emplace_back(&Sampler::samplePoint, &S, x, y, z, slot++);
        for (auto& t : threads) t.join();
        slab_ms.push_back(1e3 * (now_s() - ts));
    }
    const double total_s = now_s() - t0;
    if (slot != total) { std::fprintf(stderr, "lattice mismatch\n"); return 3; }

    std::ofstream out(argv[2], std::ios::binary);
    const int64_t cnt = (int64_t)total;
    out.write(reinterpret_cast<const char*>(&cnt), sizeof cnt);
    out.write(reinterpret_cast<const char*>(f_all.data()), (std::streamsize)(total * sizeof(double)));
    out.write(reinterpret_cast<const char*>(v_all.data()), (std::streamsize)(total * sizeof(double)));

    double steady = 0.0;                                              // slabs after the first (cold) one
    for (size_t i = 1; i < slab_ms.size(); ++i) steady += slab_ms[i];
    std::printf("{\"n\": %d, \"R\": %.17g, \"lattice\": %zu, \"slabs\": %zu, \"threads_per_slab\": %zu, \"fit_s\": %.6f, "
                "\"total_s\": %.6f, \"first_slab_ms\": %.3f, \"steady_slab_ms\": %.3f, \"calls_per_s\": %.1f, "
                "\"us_per_call\": %.3f, \"spawn_join_only_ms_per_slab\": %.3f, \"kept\": %zu, \"hw_threads\": %u}\n",
                n, R, total, slab_ms.size(), na * na, fit_s, total_s, slab_ms[0],
                slab_ms.size() > 1 ? steady / (double)(slab_ms.size() - 1) : slab_ms[0], (double)total / total_s,
                1e6 * total_s / (double)total, 1e3 * spawn_s, S.kept_x.size(), std::thread::hardware_concurrency());
    return 0;
}
