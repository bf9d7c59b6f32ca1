// Minimal stand-in for the cpptrace API used by the BOUT++ reference
// (src/sys/boutexception.cxx:47-75, src/field/field3d.cxx:880,909): no backtraces in the oracle.
#ifndef BOUT_B200_ORACLE_CPPTRACE_STUB_HPP
#define BOUT_B200_ORACLE_CPPTRACE_STUB_HPP
#include <functional>
#include <string>
#include <vector>
namespace cpptrace {
struct stacktrace_frame {
  std::string symbol;
  std::string filename;
};
struct stacktrace {
  std::vector<stacktrace_frame> frames;
  std::string to_string() const { return {}; }
};
inline stacktrace generate_trace() { return {}; }
constexpr int stdout_fileno = 1;
constexpr int stderr_fileno = 2;
inline bool isatty(int) { return false; }
class formatter {
public:
  enum class address_mode { raw, object, none };
  enum class color_mode { always, none, automatic };
  enum class symbol_mode { full, pretty, pruned };
  formatter& addresses(address_mode) { return *this; }
  formatter& break_before_filename(bool) { return *this; }
  formatter& colors(color_mode) { return *this; }
  formatter& snippets(bool) { return *this; }
  formatter& symbols(symbol_mode) { return *this; }
  formatter& filter(std::function<bool(const stacktrace_frame&)>) { return *this; }
  std::string format(const stacktrace&) const { return {}; }
};
} // namespace cpptrace
#endif
