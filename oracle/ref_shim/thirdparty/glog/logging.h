// Stand-in for google-glog (absent from the image) -- TEST INFRASTRUCTURE for oracle/_ref.
// CHECK*/LOG(FATAL) throw std::runtime_error (the reference aborts the process); INFO/WARNING go to stderr only
// when REF_VERBOSE is set.
#ifndef REF_SHIM_GLOG_LOGGING_H_
#define REF_SHIM_GLOG_LOGGING_H_
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

extern int FLAGS_logtostderr, FLAGS_alsologtostderr, FLAGS_minloglevel, FLAGS_v;

namespace google {
enum { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev) : sev_(sev) { s_ << file << ":" << line << "] "; }
  ~LogMessage() noexcept(false) {
    if (sev_ == GLOG_FATAL) throw std::runtime_error(s_.str());
    static const bool verbose = std::getenv("REF_VERBOSE") != nullptr;
    if (verbose || sev_ == GLOG_ERROR) std::cerr << "[ref " << "IWEF"[sev_] << "] " << s_.str() << std::endl;
  }
  std::ostream& stream() { return s_; }
 private:
  std::ostringstream s_;
  int sev_;
};
struct LogVoidify { void operator&(std::ostream&) {} };
struct NullStream : std::ostream { NullStream() : std::ostream(nullptr) {} };
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
template <typename T> T* CheckNotNull(const char* file, int line, const char* what, T* p) {
  if (!p) LogMessage(file, line, GLOG_FATAL).stream() << what;
  return p;
}
}  // namespace google

#define REF_LOG_INFO ::google::LogMessage(__FILE__, __LINE__, ::google::GLOG_INFO)
#define REF_LOG_WARNING ::google::LogMessage(__FILE__, __LINE__, ::google::GLOG_WARNING)
#define REF_LOG_ERROR ::google::LogMessage(__FILE__, __LINE__, ::google::GLOG_ERROR)
#define REF_LOG_FATAL ::google::LogMessage(__FILE__, __LINE__, ::google::GLOG_FATAL)
#define LOG(sev) REF_LOG_##sev.stream()
#define LOG_IF(sev, cond) !(cond) ? (void)0 : ::google::LogVoidify() & LOG(sev)
#define LOG_EVERY_N(sev, n) LOG(sev)
#define LOG_FIRST_N(sev, n) LOG(sev)
#define DLOG(sev) LOG(sev)
#define VLOG(n) LOG_IF(INFO, FLAGS_v >= (n))
#define CHECK(cond) (cond) ? (void)0 : ::google::LogVoidify() & LOG(FATAL) << "Check failed: " #cond " "
#define REF_CHECK_OP(a, b, op) ((a)op(b)) ? (void)0 : ::google::LogVoidify() & LOG(FATAL) \
    << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) REF_CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) REF_CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) REF_CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) REF_CHECK_OP(a, b, <)
#define CHECK_GE(a, b) REF_CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) REF_CHECK_OP(a, b, >)
#define CHECK_NOTNULL(p) ::google::CheckNotNull(__FILE__, __LINE__, "'" #p "' Must be non NULL", (p))
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
#endif
