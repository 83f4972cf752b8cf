// main.cpp -- `lbm` executable: the reference's command line (src/main.cpp:154-304) in front of the B200 time step.
//   lbm [-d|--debug 0|1|2] [-c|--config file | file] [-b|--bench] [-s|--solver] [-h|--help] [-v|--version]
// Runs the grid generator, hands its grid to the LBM solver (transferGrid) and runs it, in that fixed order
// (src/main.cpp:271-288).  Errors end the program the way TERMM does (src/common/term.h:6-35): message on stderr,
// exit(code) -- the shell sees 255 for code -1 (test/run.sh:55).
#include <cstdlib>
#include <iostream>
#include <string>

#include "lbm_solver.hpp"

using namespace lbmhost;

static void usage() {
  std::cout << "lbm - B200-native lattice-Boltzmann solver (drop-in for the SFCMM/LBM time step)\n"
               "Usage: lbm [OPTION...] [config.json]\n"
               "  -d, --debug arg    debug level 0..2 (accepted for compatibility)\n"
               "  -h, --help         print this help\n"
               "  -v, --version      print version\n"
               "  -c, --config arg   configuration file (default: grid.json)\n"
               "  -b, --bench        run the synthetic benchmark (256^3 D3Q19 cube; LBM_BENCH_SIZE=S for S^3)\n"
               "  -s, --solver       run the solver only\n";
}

int main(int argc, char** argv) {
  std::string config = "grid.json";
  bool bench = false;
  for(int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if(a == "-h" || a == "--help") { usage(); return 0; }
    if(a == "-v" || a == "--version") { std::cout << "lbm_b200 0.1 (reference interface: SFCMM/LBM 0.0.2)" << std::endl; return 0; }
    if(a == "-b" || a == "--bench") { bench = true; continue; }
    if(a == "-s" || a == "--solver") continue;
    if(a == "-d" || a == "--debug") { ++i; continue; }
    if(a.rfind("--debug=", 0) == 0) continue;
    if(a == "-c" || a == "--config") {
      if(i + 1 >= argc) { std::cerr << "missing argument to " << a << std::endl; return 255; }
      config = argv[++i];
      continue;
    }
    if(a.rfind("--config=", 0) == 0) { config = a.substr(9); continue; }
    if(!a.empty() && a[0] != '-') { config = a; continue; }
    std::cerr << "Unknown option " << a << std::endl;
    return 255;
  }
  RunRecord::get().enabled = true;    // gridgen_log / lbm_log in the working directory, like the reference (run_log.hpp)
  RankInfo::use_environment() = true; // one process per GPU: rank / world size from the launcher (lbm_solver.hpp: RankInfo)
  try {
    if(bench) { // the reference's --bench is unimplemented for the LBM solver (src/lbm/solver.cpp:39-46); here it runs
      LBMSolver solver;
      solver.initBenchmark(argc, argv);
      return static_cast<int>(solver.run());
    }
    GridGenerator gridder;
    gridder.init(argc, argv, config);
    gridder.run();
    LBMSolver solver;
    solver.init(argc, argv, config);
    solver.transferGrid(gridder.grid());
    return static_cast<int>(solver.run());
  } catch(const TermError& e) {
    std::cerr << "\nRank 0 threw exit code " << e.code << "\n" << e.what() << "\n\nProgram is aborting!!\n" << std::flush;
    std::exit(e.code);
  } catch(const std::exception& e) {
    std::cerr << "\nRank 0 threw exit code -1\nError: " << e.what() << "\n\nProgram is aborting!!\n" << std::flush;
    std::exit(-1);
  }
}
