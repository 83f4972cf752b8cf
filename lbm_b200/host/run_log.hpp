// run_log.hpp -- the run logs the reference leaves in the working directory: `gridgen_log` and `lbm_log` ("output
// contract kept as is").
//
// The reference's logger (/root/reference/include/common/log.h:120-156,84-100,300-330) writes an XML file: a <root> with <meta> entries
// (number of domains, creation date, user, host, directory, command line, revision, build), one <m d="domain" >text\n</m> element per
// message with the five XML characters escaped, and a closing date.  At the end of a run it holds the timer table
// (include/common/timer.h:300-400): per group a rule, a "Group <name>" line and one line per timer -- indentation two blanks per level,
// "[pp.p%] " of the parent's time, the name padded to 50 columns, the seconds right-aligned in 20 columns with six significant digits,
// " [sec]".  Scripts that read the reference's "Computation" time from lbm_log (SURVEY.md section 8d does) read this file the same way.
//
// Differences, stated: the thousands of "No periodic connection found for cells: a and b" lines of the O(N_A N_B) pairing loop are
// not written (the pairing here is a sort, grid.hpp), nor the list of unused configuration keys, nor the generator's function profile;
// the sub-timers of "Computation" (collision, propagation, ...) do not exist for a fused time step, so "Computation" has no children.
// Failing to create the file never stops a run.
#pragma once
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

namespace lbmhost {

class RunLog {
 public:
  ~RunLog() { close(); }

  void open(const std::string& name, int argc, char** argv, int domain = 0, int no_domains = 1) {
    close();
    m_domain = domain;
    m_file.open(domain == 0 ? name : name + std::to_string(domain)); // log.h: one file per domain, the root's without suffix
    if(!m_file) return;
    std::string cmd;
    for(int i = 0; i < argc; ++i) cmd += (i ? " " : "") + std::string(argv[i] != nullptr ? argv[i] : "");
    char cwd[4096];
    char host[256] = "";
    ::gethostname(host, sizeof(host) - 1);
    const char* user = std::getenv("USER");
    m_file << "<?xml version=\"1.0\" standalone=\"yes\" ?>\n<root>\n"
           << meta("noDomains", std::to_string(no_domains)) << meta("dateCreation", date()) << meta("fileFormatVersion", "1")
           << meta("user", user != nullptr ? user : "n/a") << meta("host", host) << meta("dir", ::getcwd(cwd, sizeof(cwd)) != nullptr ? cwd : "")
           << meta("executionCommand", cmd) << meta("revision", "lbm_b200 0.2") << meta("build", "nvcc sm_100a + g++ (B200)") << std::flush;
  }

  bool is_open() const { return m_file.is_open() && m_file.good(); }

  // one message element; the reference ends every message with std::endl, which lands inside the element
  void operator()(const std::string& text) {
    if(!is_open()) return;
    m_file << "<m d=\"" << m_domain << "\" >" << escape(text) << "\n</m>\n" << std::flush;
  }

  void close() {
    if(!m_file.is_open()) return;
    m_file << meta("dateClosing", date()) << "</root>\n";
    m_file.close();
  }

  static std::string escape(const std::string& s) {
    std::string o;
    o.reserve(s.size());
    for(char c : s) {
      switch(c) {
        case '"': o += "&quot;"; break;
        case '&': o += "&amp;"; break;
        case '\'': o += "&apos;"; break;
        case '<': o += "&lt;"; break;
        case '>': o += "&gt;"; break;
        default: o += c;
      }
    }
    return o;
  }

 private:
  static std::string meta(const std::string& name, const std::string& content) {
    return "<meta name=\"" + name + "\" content=\"" + escape(content) + "\" />\n";
  }
  static std::string date() {
    const std::time_t t = std::time(nullptr);
    std::tm           tmv{};
    ::localtime_r(&t, &tmv);
    char buf[32];
    std::strftime(buf, sizeof(buf), "%Y-%m-%d %H:%M:%S", &tmv);
    return buf;
  }
  std::ofstream m_file;
  int           m_domain = 0;
};

// The reference's timer tree (src/globaltimers.h, include/common/timer.h): timers are created once, started / stopped around phases,
// and printed as a table.  A timer may also be given a time measured elsewhere (the device time of the step kernels).
class RunTimers {
 public:
  int add(const std::string& name, int parent = -1) {
    m_t.push_back(Entry{name, parent, 0.0, false, {}});
    return static_cast<int>(m_t.size()) - 1;
  }
  void start(int id) {
    m_t[id].running = true;
    m_t[id].since   = clock::now();
  }
  void stop(int id) {
    if(!m_t[id].running) return;
    m_t[id].seconds += std::chrono::duration<double>(clock::now() - m_t[id].since).count();
    m_t[id].running = false;
  }
  void   set(int id, double seconds) { m_t[id].seconds = seconds; }
  double seconds(int id) const {
    return m_t[id].seconds + (m_t[id].running ? std::chrono::duration<double>(clock::now() - m_t[id].since).count() : 0.0);
  }

  // the table of one group, a message per line like the reference's logger produces
  void display(RunLog& log, const std::string& group = "Application") const {
    log(std::string(80, '-'));
    std::ostringstream h;
    h << std::left << std::setw(50) << "Group" << std::setw(40) << group;
    log(h.str());
    for(size_t i = 0; i < m_t.size(); ++i)
      if(m_t[i].parent < 0) line(log, static_cast<int>(i), 0, -1.0);
  }

 private:
  using clock = std::chrono::steady_clock;
  struct Entry {
    std::string       name;
    int               parent;
    double            seconds;
    bool              running;
    clock::time_point since;
  };
  void line(RunLog& log, int id, int indent, double parent_time) const {
    const double t   = seconds(id);
    const double pct = parent_time < 0.0 ? 100.0 : (parent_time < 1e-300 ? 0.0 : 100.0 * t / parent_time);
    std::ostringstream name;
    name << std::string(static_cast<size_t>(indent), ' ') << "[" << std::fixed << std::setprecision(1) << std::setw(4) << std::setfill('0') << std::right
         << pct << "%] " << m_t[id].name;
    std::ostringstream o;
    o << std::left << std::setw(50) << name.str() << std::right << std::setprecision(6) << std::setw(20) << t << " [sec]";
    log(o.str());
    for(size_t i = 0; i < m_t.size(); ++i)
      if(m_t[i].parent == id) line(log, static_cast<int>(i), indent + 2, t);
  }
  std::vector<Entry> m_t;
};

// process-wide logs and the timer tree shared by the generator and the solver (the reference keeps them in globals, too)
struct RunRecord {
  RunLog    grid_log, lbm_log;
  RunTimers timers;
  bool      enabled = false; // the `lbm` executable switches the files on; library users (host_capi.cpp, the tests) leave no files behind
  int total = -1, gridTotal = -1, gridInit = -1, gridCreate = -1, gridIO = -1;
  int lbmTotal = -1, lbmInit = -1, lbmMain = -1, lbmComp = -1, lbmPost = -1, lbmIO = -1;
  RunRecord() {
    total      = timers.add("Total");
    gridTotal  = timers.add("Total run time of the grid generator", total);
    gridInit   = timers.add("Init", gridTotal);
    gridCreate = timers.add("Create the grid.", gridTotal);
    gridIO     = timers.add("Grid IO.", gridTotal);
    timers.start(total);
  }
  // the solver's timers come into being with the solver (solver.cpp:23 initTimers): the generator's log, closed before, does not list them
  void createLbmTimers() {
    if(lbmTotal >= 0) return;
    lbmTotal = timers.add("Total run time of the LBM Solver.", total);
    lbmInit  = timers.add("Initialization of the LBM solver!", lbmTotal);
    lbmMain  = timers.add("Main Loop of the LBM solver!", lbmTotal);
    lbmComp  = timers.add("Computation", lbmMain);
    lbmPost  = timers.add("Postprocessing", lbmMain);
    lbmIO    = timers.add("IO", lbmMain);
  }
  static RunRecord& get() {
    static RunRecord r;
    return r;
  }
};

} // namespace lbmhost
