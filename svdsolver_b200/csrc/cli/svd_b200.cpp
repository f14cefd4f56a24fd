// svd_b200 -- command-line driver with the reference's CLI contract (svd_cuda_2.cu:1267-1447,
// README.md:77-119), on top of include/svdb200_matrix.hpp + libsvdb200.so.
//
//   svd_b200 benchmark <step> <nsteps> <ninst> <band> [float|double] [stage1|bidiag|bidiag-many|svd]
//       positional arguments in the order the reference CODE reads them (svd_cuda_2.cu:1365-1371);
//       stdout "N = <n> | <sec> sec" per size; two-line CSV (sizes, seconds) in
//       data/b200_benchmark.csv like data/cuda_2_benchmark.csv; additive extra columns on stdout:
//       GFLOP/s by 8n^3/3 and device time per stage.
//   svd_b200 check <64|512|1024> [datadir] [float|double]
//       band fixed to 4 (svd_cuda_2.cu:1300).  Reads test_<type>_N_N.bin, band_<type>_N_N.bin,
//       bidiagonal_<type>_N_N.bin from datadir (the reference hard-codes /data/spencerrose/), runs
//       stage 1 in panel order (what cuda_brd_p1 is) and in tile order (what produced the fixture),
//       then stage 2, and prints the reference's MSE metric (matrix_gpu.h:438) plus the signed
//       max-abs-diff over the band diagonals.  Exit code 0 always, like the reference.
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <vector>

#include "svdb200_matrix.hpp"

namespace {

void print_help() {
    std::cout << "Options for B200 SVD testing" << std::endl;
    std::cout << "\n(1) Run benchmark tests for the band / bidiagonal reduction." << std::endl;
    std::cout << "\t>> benchmark [<int> Step size] [<int> Number of steps] [<int> Number of test instances] [<int> Band size] [float|double] [stage1|bidiag|bidiag-many|svd]";
    std::cout << "\n\tExample: ./svd_b200 benchmark 320 12 1 32" << std::endl;
    std::cout << "\n(2) Correctness Test: compares test matrix and corresponding band and bidiagonal reductions" << std::endl;
    std::cout << "\t>> check [64|512|1024 Row/Column sizes] [data dir] [float|double]" << std::endl;
    std::cout << "\tExample: ./svd_b200 check 64 tests/golden\n" << std::endl;
}

// matrix_generator (svd_cuda_2.cu:1230-1245) with a reproducible stream (seed 586 + n + instance)
template <typename T>
std::vector<csc586::gpu::Matrix<T>> generate(size_t rows, size_t cols, size_t count, T lo, T hi) {
    std::vector<csc586::gpu::Matrix<T>> v;
    for (size_t i = 0; i < count; ++i) {
        csc586::gpu::Matrix<T> m(rows, cols);
        m.fill(lo, hi, 586 + rows + i);
        v.push_back(m);
    }
    return v;
}

template <typename T>
double band_rel(const csc586::Matrix<T>& a, const csc586::Matrix<T>& ref, size_t band) {
    double num = 0, den = 0;
    for (size_t i = 0; i < ref.nrows; ++i)
        for (size_t j = 0; j < ref.ncols; ++j) {
            den = std::max(den, std::abs((double)ref[i][j]));
            if (j >= i && j <= i + band) num = std::max(num, std::abs((double)a[i][j] - (double)ref[i][j]));
        }
    return den > 0 ? num / den : num;
}

template <typename T>
int run_check(const std::string& nstr, const std::string& dir, const char* tname) {
    std::cout << "Checking correctness ... " << std::endl;
    const size_t size = (size_t)std::atoi(nstr.c_str());
    const size_t band_size = 4u;
    auto path = [&](const char* kind) { return dir + "/" + kind + "_" + tname + "_" + nstr + "_" + nstr + ".bin"; };
    csc586::gpu::Matrix<T> A(size, size), band_check(size, size), brd_check(size, size);
    std::cout << "Reading file: " << path("test") << std::endl;
    if (!A.read(path("test")) || !band_check.read(path("band")) || !brd_check.read(path("bidiagonal"))) {
        std::cout << "fixtures not found in " << dir << std::endl;
        return 0;
    }
    A.print();
    auto A_panel = A;
    std::cout << "\n\nB200 Test (Band, panel order = cuda_brd_p1):" << std::endl;
    csc586::gpu::cuda_brd_p1(A_panel, band_size);
    A_panel.print(16);
    std::cout << "\n\nBaseline Test (Band):" << std::endl;
    band_check.print(16);
    std::cout << "\n\nMSE of Band Reduction: " << A_panel.mse(band_check, band_size) << std::endl;
    auto A_tile = A;
    csc586::parallel::brd_p1(A_tile, band_size);
    std::cout << "MSE of Band Reduction (tile order = parallel::brd_p1): " << A_tile.mse(band_check, band_size)
              << "   signed rel diff over diagonals 0.." << band_size << ": " << band_rel(A_tile, band_check, band_size) << std::endl;
    std::cout << "\n\nB200 Test (Bidiagonal, brd_p2 on the tile-order band):" << std::endl;
    auto B = csc586::parallel::brd_p2(A_tile, band_size);
    A_tile.print(10);
    std::cout << "\n\nBaseline Test (Bidiagonal):" << std::endl;
    brd_check.print(10);
    std::cout << "\n\nMSE of Bidiagonal Reduction: " << A_tile.mse(brd_check, 2)
              << "   signed rel diff over diagonals 0..1: " << band_rel(A_tile, brd_check, 1) << std::endl;
    {   // the list entry point against one call per instance (panel order)
        std::vector<csc586::gpu::Matrix<T>> many{A, A, A};
        auto Bm = csc586::gpu::cuda_bidiagonalize_many<T>(many, band_size);
        auto A_one = A;
        csc586::gpu::cuda_brd_p1(A_one, band_size);
        auto B1 = csc586::parallel::brd_p2(A_one, band_size);
        double md = 0;
        for (auto& Bi : Bm) {
            for (size_t i = 0; i < B1.d.size(); ++i) md = std::max(md, std::abs((double)Bi.d[i] - (double)B1.d[i]));
            for (size_t i = 0; i < B1.e.size(); ++i) md = std::max(md, std::abs((double)Bi.e[i] - (double)B1.e[i]));
        }
        std::cout << "Max |diff| cuda_bidiagonalize_many vs one call per instance: " << md << std::endl;
    }
    auto sig = csc586::serial::qrd(B);
    std::cout << "Largest / smallest singular value (QR diagonalisation): " << sig.d.front() << " / " << sig.d.back() << std::endl;
    return 0;
}

template <typename T>
int run_benchmark(int argc, char* argv[], const std::string& what) {
    const T min_val = 0, max_val = 5;                     // svd_cuda_2.cu:1361-1362
    auto step = size_t(std::atoi(argv[2]));
    auto nsteps = size_t(std::atoi(argv[3]) + 1);
    auto n_test_instances = size_t(std::atoi(argv[4]));
    auto b_size = size_t(std::atoi(argv[5]));
    (void)argc;
    std::vector<int> x;
    std::vector<float> y;
    std::cout << "Benchmark: B200 " << (what == "stage1" ? "Band Reduction" : what == "bidiag" ? "Bidiagonal Reduction" : what == "bidiag-many" ? "Bidiagonal Reduction (instances handed over together)" : "Singular Values") << std::endl;
    std::cout << "\tBand size: " << b_size << std::endl;
    std::cout << "\tStep size: " << step << std::endl;
    std::cout << "\tNumber of steps: " << nsteps - 1 << std::endl;
    std::cout << "\tNumber of test instances: " << n_test_instances << std::endl;
    csc586::gpu::Matrix<T> (*brd_p1)(csc586::gpu::Matrix<T>&, const size_t) = csc586::gpu::cuda_brd_p1<T>;
    auto bidiag = [](csc586::gpu::Matrix<T>& a, const size_t b) { csc586::gpu::cuda_brd_p1<T>(a, b); return csc586::parallel::brd_p2<T>(a, b); };
    auto svd = [](csc586::gpu::Matrix<T>& a, const size_t b) {
        csc586::gpu::cuda_brd_p1<T>(a, b);
        auto B = csc586::parallel::brd_p2<T>(a, b);
        return csc586::serial::qrd<T>(B);
    };
    {   // warm-up: context + workspace creation (sized for the largest matrix of the run) are not part of the per-instance time
        auto w = generate<T>(step, step, 2, min_val, max_val);
        csc586::b200::session<T>().get((nsteps - 1) * step, b_size, csc586::b200::dtype_code<T>());
        brd_p1(w[0], b_size);
        if (what == "bidiag-many") csc586::gpu::cuda_bidiagonalize_many<T>(w, b_size);
    }
    std::cout << "Average time per B200 reduction" << std::endl;
    for (size_t k = 1; k < nsteps; ++k) {
        const size_t rows = k * step, cols = k * step;
        auto data = generate<T>(rows, cols, n_test_instances, min_val, max_val);
        float avg = 0;
        if (what == "stage1") avg = csc586::benchmark::benchmark(brd_p1, data, b_size);
        else if (what == "bidiag") avg = csc586::benchmark::benchmark(bidiag, data, b_size);
        else if (what == "bidiag-many") avg = csc586::benchmark::benchmark_many(csc586::gpu::cuda_bidiagonalize_many<T>, data, b_size);
        else avg = csc586::benchmark::benchmark(svd, data, b_size);
        const double sec = avg * 1e-6;
        std::cout << "N = " << cols << " | " << sec << " sec"
                  << " | " << (8.0 * cols * cols * cols / 3.0) / sec * 1e-9 << " GFLOP/s (8n^3/3, wall clock incl. H2D/D2H)" << std::endl;
        x.push_back((int)(k * step));
        y.push_back((float)sec);
    }
    std::ostringstream vts;
    if (!x.empty()) {
        std::copy(x.begin(), x.end() - 1, std::ostream_iterator<int>(vts, ", "));
        vts << x.back() << "\n";
        std::copy(y.begin(), y.end() - 1, std::ostream_iterator<float>(vts, ", "));
        vts << y.back();
    }
    auto filename = std::string("data/b200_benchmark.csv");
    std::cout << "Writing results to file ... " << filename << std::endl;
    std::ofstream ftest(filename);
    ftest << vts.str();
    ftest.close();
    std::cout << "Done." << std::endl;
    return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
    if (argc >= 3 && std::strncmp(argv[1], "check", 5) == 0) {
        std::string dir = argc > 3 ? argv[3] : "tests/golden";
        bool dbl = argc > 4 && std::strcmp(argv[4], "double") == 0;
        return dbl ? run_check<double>(argv[2], dir, "double") : run_check<float>(argv[2], dir, "float");
    }
    if (argc > 5 && std::strncmp(argv[1], "benchmark", 9) == 0) {
        bool dbl = argc > 6 && std::strcmp(argv[6], "double") == 0;
        std::string what = argc > 7 ? argv[7] : "stage1";
        return dbl ? run_benchmark<double>(argc, argv, what) : run_benchmark<float>(argc, argv, what);
    }
    print_help();
    return 0;
}
