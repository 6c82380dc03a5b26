// miniapp_cholesky — benchmark driver with the command line and the output lines of the reference's
// miniapp/miniapp_cholesky.cpp (options :217-231 + miniapp/include/dlaf/miniapp/options.h:237-263, run loop
// :107-198, output :165-188, check :408-446), written against this repository's C++ surface
// (include/dlaf/...). No MPI / pika: ranks are separate processes (one per GPU) that find each other through
// RANK / WORLD_SIZE / LOCAL_RANK (torchrun-compatible) and exchange the NCCL id through a file.
//
//   miniapp_cholesky --matrix-size 32768 --block-size 512 --type d --uplo L --nruns 3 --nwarmups 1
//                    --grid-rows 1 --grid-cols 1 --check-result last [--csv] [--dlaf:print-config]
#include <unistd.h>

#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include <dlaf/communication/communicator_grid.h>
#include <dlaf/factorization/cholesky.h>
#include <dlaf/init.h>
#include <dlaf/matrix/matrix.h>
#include <dlaf/types.h>
#include <dlaf/util_matrix.h>
#include <dlaf_c/b200_ext.h>
#include <dlaf_c/grid.h>

using namespace dlaf;

struct Options {
  SizeType m = 4096, mb = 256;  // miniapp_cholesky.cpp:223-224
  int grid_rows = 1, grid_cols = 1;
  bool local = false;
  int64_t nruns = 1, nwarmups = 1;
  std::string check = "none", type = "d", backend = "default", uplo = "L", info;
  bool csv = false;
  std::vector<std::string> dlaf_args;
};

[[noreturn]] static void invalid(const std::string& opt, const std::string& got, const std::string& expected) {
  std::cout << "Invalid option for " << opt << ". Got '" << got << "' but expected one of " << expected << "."
            << std::endl;
  std::terminate();
}

static Options parse(int argc, char** argv) {
  Options o;
  auto value = [&](int& i, const std::string& arg, const std::string& name) -> std::string {
    const std::string eq = "--" + name + "=";
    if (arg.rfind(eq, 0) == 0)
      return arg.substr(eq.size());
    if (i + 1 >= argc) {
      std::cout << "missing value for --" << name << std::endl;
      std::terminate();
    }
    return argv[++i];
  };
  auto is = [](const std::string& arg, const std::string& name) {
    return arg == "--" + name || arg.rfind("--" + name + "=", 0) == 0;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.rfind("--dlaf:", 0) == 0) o.dlaf_args.push_back(a);
    else if (a.rfind("--pika:", 0) == 0) continue;  // accepted and ignored: there is no pika runtime
    else if (is(a, "matrix-size")) o.m = std::stoll(value(i, a, "matrix-size"));
    else if (is(a, "block-size")) o.mb = std::stoll(value(i, a, "block-size"));
    else if (is(a, "grid-rows")) o.grid_rows = std::stoi(value(i, a, "grid-rows"));
    else if (is(a, "grid-cols")) o.grid_cols = std::stoi(value(i, a, "grid-cols"));
    else if (a == "--local") o.local = true;
    else if (is(a, "nruns")) o.nruns = std::stoll(value(i, a, "nruns"));
    else if (is(a, "nwarmups")) o.nwarmups = std::stoll(value(i, a, "nwarmups"));
    else if (is(a, "check-result")) o.check = value(i, a, "check-result");
    else if (a == "--csv") o.csv = true;
    else if (is(a, "pp-info")) o.info = value(i, a, "pp-info");
    else if (is(a, "type")) o.type = value(i, a, "type");
    else if (is(a, "backend")) o.backend = value(i, a, "backend");
    else if (is(a, "uplo")) o.uplo = value(i, a, "uplo");
    else if (a == "--help" || a == "-h") {
      std::cout << "Allowed options: --matrix-size --block-size --grid-rows --grid-cols --local --nruns --nwarmups\n"
                   "  --check-result none|last|all --csv --pp-info --type s|d|c|z --backend default|gpu --uplo L|U --dlaf:*\n";
      std::exit(0);
    }
    else {
      std::cout << "unrecognised option '" << a << "'" << std::endl;
      std::terminate();
    }
  }
  for (auto& ch : o.type) ch = static_cast<char>(std::tolower(ch));
  for (auto& ch : o.uplo) ch = static_cast<char>(std::toupper(ch));
  if (o.type.size() != 1 || std::string("sdcz").find(o.type) == std::string::npos) invalid("--type", o.type, "'s', 'd', 'c', 'z'");
  if (o.uplo != "L" && o.uplo != "U") invalid("--uplo", o.uplo, "'L', 'U'");
  if (o.check != "none" && o.check != "last" && o.check != "all") invalid("--check-result", o.check, "'none', 'last', 'all'");
  if (o.backend == "mc") {
    std::cout << "Asked for --backend=mc but this build has the GPU backend only (no CPU path)." << std::endl;
    std::terminate();
  }
  if (o.backend != "default" && o.backend != "gpu") invalid("--backend", o.backend, "'default', 'gpu'");
  if (o.local && (o.grid_rows != 1 || o.grid_cols != 1)) {  // options.h:226
    std::cout << "--local requires a 1x1 grid" << std::endl;
    std::terminate();
  }
  return o;
}

// NCCL id rendezvous between the ranks of one node through a file (replaces MPI_Init + MPI_COMM_WORLD).
static DLAF_Comm bootstrap(int rank, int size) {
  if (size == 1)
    return nullptr;
  const char* port = std::getenv("MASTER_PORT");
  const char* custom = std::getenv("DLAF_B200_RENDEZVOUS");
  const std::string path = custom ? custom : std::string("/tmp/dlaf_b200_id_") + (port ? port : "0");
  char id[DLAF_B200_UNIQUE_ID_BYTES];
  if (rank == 0) {
    dlaf_b200_get_unique_id(id);
    std::ofstream f(path + ".tmp", std::ios::binary);
    f.write(id, sizeof(id));
    f.close();
    std::rename((path + ".tmp").c_str(), path.c_str());
  }
  else {
    for (int tries = 0;; ++tries) {
      std::ifstream f(path, std::ios::binary);
      if (f && f.read(id, sizeof(id)))
        break;
      if (tries > 6000) {
        std::cerr << "rendezvous file " << path << " never appeared" << std::endl;
        std::terminate();
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
  }
  DLAF_Comm c = dlaf_b200_comm_create(id, rank, size);
  if (rank == 0)
    std::remove(path.c_str());
  return c;
}

static double check_call(int ctx, char uplo, const float* a, const float* f, DLAF_descriptor d) { return dlaf_b200_check_cholesky_s(ctx, uplo, a, f, d); }
static double check_call(int ctx, char uplo, const double* a, const double* f, DLAF_descriptor d) { return dlaf_b200_check_cholesky_d(ctx, uplo, a, f, d); }
static double check_call(int ctx, char uplo, const std::complex<float>* a, const std::complex<float>* f, DLAF_descriptor d) { return dlaf_b200_check_cholesky_c(ctx, uplo, a, f, d); }
static double check_call(int ctx, char uplo, const std::complex<double>* a, const std::complex<double>* f, DLAF_descriptor d) { return dlaf_b200_check_cholesky_z(ctx, uplo, a, f, d); }

template <class T>
static void run(const Options& opts, DLAF_Comm world, int world_rank) {
  constexpr Backend backend = Backend::GPU;
  using HostMatrix = Matrix<T, Device::CPU>;
  using MirrorType = matrix::MatrixMirror<T, DefaultDevice_v<backend>, Device::CPU>;
  comm::CommunicatorGrid comm_grid(world, opts.grid_rows, opts.grid_cols, common::Ordering::ColumnMajor);  // :113
  const GlobalElementSize matrix_size(opts.m, opts.m);
  const TileElementSize block_size(opts.mb, opts.mb);
  const blas::Uplo uplo = opts.uplo == "L" ? blas::Uplo::Lower : blas::Uplo::Upper;

  HostMatrix matrix_ref(matrix_size, block_size, comm_grid);
  matrix::util::set_random_hermitian_positive_definite(matrix_ref);

  for (int64_t run_index = -opts.nwarmups; run_index < opts.nruns; ++run_index) {
    if (world_rank == 0 && run_index >= 0)
      std::cout << "[" << run_index << "]" << std::endl;
    HostMatrix matrix_host(matrix_size, block_size, comm_grid);
    matrix::copy(matrix_ref, matrix_host);
    matrix_host.waitLocalTiles();
    double elapsed_time;
    {
      MirrorType matrix(matrix_host, comm_grid);
      matrix.get().waitLocalTiles();       // input resident on the device
      comm_grid.wait_all_communicators();  // MPI_Barrier of the reference (:141)
      const auto t0 = std::chrono::steady_clock::now();
      if (opts.local)
        cholesky_factorization<backend, DefaultDevice_v<backend>, T>(uplo, matrix.get());
      else
        cholesky_factorization<backend, DefaultDevice_v<backend>, T>(comm_grid, uplo, matrix.get());
      const int info = cholesky_info(matrix.get());  // waitLocalTiles + info
      comm_grid.wait_all_communicators();
      elapsed_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (info != 0 && world_rank == 0)
        std::cout << "ERROR: the matrix is not positive definite (info = " << info << ")" << std::endl;
    }
    const double n = static_cast<double>(opts.m);
    const double add_mul = n * n * n / 6;
    const double gigaflops = total_ops<T>(add_mul, add_mul) / elapsed_time / 1e9;
    if (world_rank == 0 && run_index >= 0) {
      std::cout << "[" << run_index << "]"
                << " " << elapsed_time << "s"
                << " " << gigaflops << "GFlop/s"
                << " " << opts.type << opts.uplo << " " << matrix_size << " " << block_size << " "
                << comm_grid.size() << " " << 1 << " "
                << "GPU" << std::endl;
      if (opts.csv)
        std::cout << "CSVData-2, "
                  << "run, " << run_index << ", "
                  << "time, " << elapsed_time << ", "
                  << "GFlops, " << gigaflops << ", "
                  << "type, " << opts.type << ", "
                  << "UpLo, " << opts.uplo << ", "
                  << "matrixsize, " << opts.m << ", "
                  << "blocksize, " << opts.mb << ", "
                  << "comm_rows, " << opts.grid_rows << ", "
                  << "comm_cols, " << opts.grid_cols << ", "
                  << "threads, " << 1 << ", "
                  << "backend, " << "GPU" << ", " << opts.info << std::endl;
    }
    if ((opts.check == "last" && run_index == opts.nruns - 1) || opts.check == "all") {
      const double ratio = check_call(comm_grid.context(), opts.uplo[0], matrix_ref.ptr(), matrix_host.ptr(),
                                      matrix_host.descriptor());
      if (world_rank == 0) {
        // collective over the grid (every rank made the call above); same gate as miniapp_cholesky.cpp:436-445
        const double eps = std::numeric_limits<BaseType<T>>::epsilon();
        if (ratio > 100 * eps * n)
          std::cout << "ERROR: ";
        else if (ratio > eps * n)
          std::cout << "Warning: ";
        std::cout << "Max Diff / Max A: " << ratio << std::endl;
      }
    }
  }
}

int main(int argc, char** argv) {
  const Options opts = parse(argc, argv);
  const int rank = std::getenv("RANK") ? std::atoi(std::getenv("RANK")) : 0;
  const int size = std::getenv("WORLD_SIZE") ? std::atoi(std::getenv("WORLD_SIZE")) : 1;
  std::vector<const char*> dargs;
  dargs.push_back("miniapp_cholesky");
  for (const auto& s : opts.dlaf_args)
    dargs.push_back(s.c_str());
  ScopedInitializer init(static_cast<int>(dargs.size()), dargs.data());
  if (opts.grid_rows * opts.grid_cols > size) {
    std::cout << "grid " << opts.grid_rows << "x" << opts.grid_cols << " needs " << opts.grid_rows * opts.grid_cols
              << " ranks, launched with " << size << std::endl;
    std::terminate();
  }
  DLAF_Comm world = bootstrap(rank, size);
  switch (opts.type[0]) {
    case 's': run<float>(opts, world, rank); break;
    case 'd': run<double>(opts, world, rank); break;
    case 'c': run<std::complex<float>>(opts, world, rank); break;
    default: run<std::complex<double>>(opts, world, rank); break;
  }
  if (world)
    dlaf_b200_comm_destroy(world);
  return 0;
}
