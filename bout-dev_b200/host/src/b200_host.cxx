// Implementation of the C++17 host layer (see include/bout/b200_host.hxx) on the C-ABI of
// include/bout_b200.h.  Set-up work (options, metric, initial profiles) runs once on the host, as in
// the reference; everything inside a right-hand side or a time step is a kernel launch.
#include "bout/b200_host.hxx"
#include "bout/field_accessor.hxx"

#include <deque>
#include <thread>

#include <algorithm>
#include <cctype>
#include <cstring>
#include <stdexcept>

Output output, output_info, output_warn, output_progress, output_error;

namespace bout {
namespace globals {
Mesh* mesh = nullptr;
}
namespace b200 {

void check(int status) {
  if (status != BOUT_B200_OK) {
    throw BoutException("bout_b200: {} (status {})", bout_b200_last_error(), status);
  }
}

// ---- device block pool (reference Array<T> arena, include/bout/array.hxx:387-472) -------------------
namespace {
std::map<size_t, std::vector<double*>>& freeList() {
  static std::map<size_t, std::vector<double*>> pool;
  return pool;
}
size_t n_malloc = 0;
} // namespace

void Block::syncToHost() {
  if (mirror == none) {
    host.resize(count);
    check(bout_b200_memcpy_d2h(host.data(), ptr, count * sizeof(double), nullptr));
    mirror = synced;
  }
}
void Block::syncToDevice() {
  if (mirror == host_newer) {
    check(bout_b200_memcpy_h2d(ptr, host.data(), count * sizeof(double), nullptr));
    mirror = synced;
  }
}

void Block::ensure() {
  if (pending.empty()) {
    return;
  }
  std::vector<LazyTerm> terms;
  terms.swap(pending); // (an operand can never be this block: it was created after all of them)
  std::vector<bout_b200_ewterm> t(terms.size());
  auto devptr = [](const std::shared_ptr<Block>& b) -> const double* {
    if (!b) {
      return nullptr;
    }
    b->ensure();
    b->syncToDevice();
    return b->ptr;
  };
  for (size_t i = 0; i < terms.size(); ++i) {
    t[i] = bout_b200_ewterm{};
    t[i].coef = terms[i].coef;
    t[i].f1 = devptr(terms[i].f1);
    t[i].f2 = devptr(terms[i].f2);
    t[i].div = devptr(terms[i].dv);
    t[i].f1_is2d = terms[i].f1_2d;
    t[i].f2_is2d = terms[i].f2_2d;
    t[i].div_is2d = terms[i].dv_2d;
  }
  const bout_b200_grid g{pgrid[0], pgrid[1], pgrid[2], pgrid[3], pgrid[4], pgrid[5], pgrid[6]};
  check(bout_b200_elementwise(&g, BOUT_B200_RGN_ALL, static_cast<int>(t.size()), t.data(), ptr, nullptr));
  mirror = none;
}

void Block::launchDependents() {
  if (dependents.empty()) {
    return;
  }
  std::vector<std::weak_ptr<Block>> deps;
  deps.swap(dependents);
  for (auto& w : deps) {
    if (auto d = w.lock()) {
      d->ensure();
    }
  }
}

Block::~Block() {
  if (ptr != nullptr) {
    freeList()[count].push_back(ptr); // back to the pool, not to the driver
  }
}

StageFusion& stageFusion() {
  static StageFusion f;
  return f;
}

unsigned long long nextVersion() {
  static unsigned long long v = 0;
  return ++v;
}

std::shared_ptr<Block> acquire(size_t count) {
  auto blk = std::make_shared<Block>();
  blk->count = count;
  blk->version = nextVersion();
  auto& bucket = freeList()[count];
  if (!bucket.empty()) {
    blk->ptr = bucket.back();
    bucket.pop_back();
  } else {
    void* p = nullptr;
    check(bout_b200_malloc(&p, count * sizeof(double)));
    blk->ptr = static_cast<double*>(p);
    ++n_malloc;
  }
  return blk;
}

void releasePool() {
  for (auto& kv : freeList()) {
    for (double* p : kv.second) {
      bout_b200_free(p);
    }
  }
  freeList().clear();
}

size_t poolAllocations() { return n_malloc; }
static size_t n_cow = 0;
size_t cowCopies() { return n_cow; }
void countCowCopy() { ++n_cow; }

// ---- small expression evaluator for option values and initial profiles ---------------------------------
// (the reference's ExpressionParser / FieldFactory, src/sys/expressionparser.cxx, src/field/fieldgenerators.cxx,
//  restated for the subset the example inputs use: + - * / ^ ( ) numbers, pi, x y z t, option references,
//  sin cos tan exp log sqrt abs gauss mixmode tanh H)
struct EvalContext {
  double x = 0, y = 0, z = 0, t = 0;
  Options* section = nullptr;
  int depth = 0;
};

class Expr {
public:
  Expr(const std::string& s, const EvalContext& c) : src(s), ctx(c) {}
  double eval() {
    pos = 0;
    const double v = sum();
    skip();
    if (pos != src.size()) {
      throw BoutException("Cannot parse expression '{}' near position {}", src, pos);
    }
    return v;
  }

private:
  std::string src;
  EvalContext ctx;
  size_t pos = 0;
  void skip() {
    while (pos < src.size() && std::isspace(static_cast<unsigned char>(src[pos]))) {
      ++pos;
    }
  }
  bool eat(char c) {
    skip();
    if (pos < src.size() && src[pos] == c) {
      ++pos;
      return true;
    }
    return false;
  }
  // Binary operators by precedence climbing with the reference's table and quirks (src/sys/expressionparser.cxx:
  // 256-264 precedences, :554-608 parseBinOpRHS, all operators left-associative; :530-534 a '-' where an operand is
  // expected yields an implicit 0 and stays a BINARY minus, so "a * -b" is (a * 0) - b as in the reference;
  // :682-687 a number directly followed by a symbol multiplies it, "2π").
  bool after_number = false;
  static bool identStart(char c) {
    return std::isalpha(static_cast<unsigned char>(c)) || c == '_' || static_cast<unsigned char>(c) >= 0x80;
  }
  static int precedence(char op) {
    switch (op) {
    case '|': return 3;
    case '&': return 5;
    case '<':
    case '>': return 7;
    case '+':
    case '-': return 10;
    case '*':
    case '/': return 20;
    case '^': return 30;
    default: return -1;
    }
  }
  static bool toBool(double v, const std::string& where) { // expressionparser.cxx:196-206
    if (std::fabs(v - 1.0) < 1e-5) {
      return true;
    }
    if (std::fabs(v) < 1e-5) {
      return false;
    }
    throw BoutException("Boolean operator argument {} is not a bool in '{}'", v, where);
  }
  // next binary operator without consuming it: 0 = none; 'i' = the implicit '*' after a number
  char peekOp() {
    skip();
    if (pos >= src.size()) {
      return 0;
    }
    const char c = src[pos];
    if (after_number && identStart(c)) {
      return 'i';
    }
    return precedence(c) >= 0 ? c : 0;
  }
  double apply(char op, double l, double r) const {
    switch (op) {
    case '|': return (toBool(l, src) || toBool(r, src)) ? 1.0 : 0.0;
    case '&': return (toBool(l, src) && toBool(r, src)) ? 1.0 : 0.0;
    case '>': return l > r ? 1.0 : 0.0;
    case '<': return l < r ? 1.0 : 0.0;
    case '+': return l + r;
    case '-': return l - r;
    case '/': return l / r;
    case '^': return std::pow(l, r);
    default: return l * r; // '*' and the implicit product
    }
  }
  double binopRHS(int exprPrec, double lhs) {
    for (;;) {
      const char op = peekOp();
      if (op == 0) {
        return lhs;
      }
      const int prec = precedence(op == 'i' ? '*' : op);
      if (prec < exprPrec) {
        return lhs;
      }
      if (op != 'i') {
        ++pos;
      }
      double rhs = atom();
      const char next = peekOp();
      if (next != 0 && prec < precedence(next == 'i' ? '*' : next)) {
        rhs = binopRHS(prec + 1, rhs);
      }
      lhs = apply(op, lhs, rhs);
    }
  }
  double sum() { // a whole expression
    const double lhs = atom();
    return binopRHS(0, lhs);
  }
  static double genRand(double seed) { // FieldMixmode::genRand, fieldgenerators.cxx:148-167
    if (seed < 0.0) {
      seed *= -1;
    }
    const int niter = 11 + (23 + static_cast<int>(std::lround(seed))) % 79;
    const double A = 0.01, B = 1.23456789;
    double x = (A + std::fmod(seed, B)) / (B + 2. * A);
    for (int i = 0; i != niter; ++i) {
      x = 3.99 * x * (1. - x);
    }
    return x;
  }
  double call(const std::string& fn, const std::vector<double>& a) {
    auto need = [&](size_t lo, size_t hi) {
      if (a.size() < lo || a.size() > hi) {
        throw BoutException("Wrong number of arguments to {} in '{}'", fn, src);
      }
    };
    if (fn == "sin") { need(1, 1); return std::sin(a[0]); }
    if (fn == "cos") { need(1, 1); return std::cos(a[0]); }
    if (fn == "tan") { need(1, 1); return std::tan(a[0]); }
    if (fn == "exp") { need(1, 1); return std::exp(a[0]); }
    if (fn == "log") { need(1, 1); return std::log(a[0]); }
    if (fn == "sqrt") { need(1, 1); return std::sqrt(a[0]); }
    if (fn == "abs") { need(1, 1); return std::fabs(a[0]); }
    if (fn == "tanh") { need(1, 1); return std::tanh(a[0]); }
    if (fn == "h" || fn == "H") { need(1, 1); return a[0] > 0.0 ? 1.0 : 0.0; }
    if (fn == "gauss") { // FieldGaussian, fieldgenerators.cxx:31-34
      need(1, 2);
      const double sigma = a.size() == 2 ? a[1] : 1.0;
      return std::exp(-SQ(a[0] / sigma) / 2.) / (std::sqrt(TWOPI) * sigma);
    }
    if (fn == "mixmode") { // FieldMixmode, fieldgenerators.cxx:110-146
      need(1, 2);
      const double seed = a.size() == 2 ? a[1] : 0.5;
      // the 14 phases depend on the seed only: kept from the previous call (a field is filled point by point)
      thread_local double cached_seed = std::nan("");
      thread_local double phase[14];
      if (!(cached_seed == seed)) {
        for (int i = 0; i < 14; i++) {
          phase[i] = PI * (2. * genRand(seed + i) - 1.);
        }
        cached_seed = seed;
      }
      double result = 0.0;
      for (int i = 0; i < 14; i++) {
        result += (1. / SQ(1. + std::abs(i - 4))) * std::cos(i * a[0] + phase[i]);
      }
      return result;
    }
    if (fn == "min") { need(1, 8); return *std::min_element(a.begin(), a.end()); }
    if (fn == "max") { need(1, 8); return *std::max_element(a.begin(), a.end()); }
    if (fn == "power") { need(2, 2); return std::pow(a[0], a[1]); }
    if (fn == "sinh") { need(1, 1); return std::sinh(a[0]); }
    if (fn == "cosh") { need(1, 1); return std::cosh(a[0]); }
    if (fn == "asin") { need(1, 1); return std::asin(a[0]); }
    if (fn == "acos") { need(1, 1); return std::acos(a[0]); }
    if (fn == "atan") { need(1, 2); return a.size() == 1 ? std::atan(a[0]) : std::atan2(a[0], a[1]); } // fieldgenerators.hxx:103-108
    if (fn == "erf") { need(1, 1); return std::erf(a[0]); }
    if (fn == "fmod") { need(2, 2); return std::fmod(a[0], a[1]); }
    if (fn == "clamp") { need(3, 3); return a[0] < a[1] ? a[1] : (a[0] > a[2] ? a[2] : a[0]); } // fieldgenerators.hxx:219-231
    if (fn == "round") { need(1, 1); return static_cast<int>(a[0] > 0.0 ? a[0] + 0.5 : a[0] - 0.5); } // :249-255
    if (fn == "tanhhat") { // FieldTanhHat (X, width, centre, steepness), fieldgenerators.cxx:198-206
      need(4, 4);
      return 0.5 * (std::tanh(a[3] * (a[0] - (a[2] - 0.5 * a[1]))) - std::tanh(a[3] * (a[0] - (a[2] + 0.5 * a[1]))));
    }
    if (fn == "where") { need(3, 3); return a[0] > 0.0 ? a[1] : a[2]; } // FieldWhere, fieldgenerators.cxx:208-214
    throw BoutException("Unknown function '{}' in expression '{}'", fn, src);
  }
  double lookup(const std::string& id) {
    if (id == "pi" || id == "π") { return PI; }
    if (id == "true" || id == "True") { return 1.0; } // field_factory.cxx:120-123
    if (id == "false" || id == "False") { return 0.0; }
    if (id == "x") { return ctx.x; }
    if (id == "y") { return ctx.y; }
    if (id == "z") { return ctx.z; }
    if (id == "t") { return ctx.t; }
    // option reference: same section, then parents, "section:key" from the root
    if (ctx.depth > 16) {
      throw BoutException("Option reference loop while evaluating '{}'", src);
    }
    Options* found = nullptr;
    const auto colon = id.find(':');
    if (colon != std::string::npos) {
      Options* o = &Options::root();
      size_t start = 0;
      for (;;) {
        const auto c = id.find(':', start);
        const std::string part = id.substr(start, c == std::string::npos ? std::string::npos : c - start);
        if (c == std::string::npos) {
          if (o->isSet(part)) {
            found = &(*o)[part];
          }
          break;
        }
        o = &(*o)[part];
        start = c + 1;
      }
    } else {
      for (Options* o = ctx.section; o != nullptr; o = o->parent()) {
        if (o->isSet(id)) {
          found = &(*o)[id];
          break;
        }
      }
    }
    if (found == nullptr) {
      throw BoutException("Unknown symbol '{}' in expression '{}'", id, src);
    }
    EvalContext c = ctx;
    c.section = found->parent();
    c.depth = ctx.depth + 1;
    return Expr(found->text(), c).eval();
  }
  double atom() {
    skip();
    after_number = false;
    if (pos >= src.size()) {
      throw BoutException("Unexpected end of expression '{}'", src);
    }
    if (src[pos] == '-' || src[pos] == '+') {
      return 0.0; // implicit zero; the sign is then read as a binary operator (expressionparser.cxx:530-534)
    }
    if (eat('!')) {
      return toBool(atom(), src) ? 0.0 : 1.0;
    }
    if (eat('(')) {
      const double v = sum();
      if (!eat(')')) {
        throw BoutException("Missing ')' in '{}'", src);
      }
      return v;
    }
    const char c = src[pos];
    if (std::isdigit(static_cast<unsigned char>(c)) || c == '.') {
      size_t used = 0;
      const double v = std::stod(src.substr(pos), &used);
      pos += used;
      after_number = true;
      return v;
    }
    if (identStart(c)) { // symbols may hold any non-reserved character, UTF-8 included (expressionparser.cxx:679-720)
      size_t e = pos;
      while (e < src.size() && (std::isalnum(static_cast<unsigned char>(src[e])) || src[e] == '_' || src[e] == ':'
                                || static_cast<unsigned char>(src[e]) >= 0x80)) {
        ++e;
      }
      const std::string id = src.substr(pos, e - pos);
      pos = e;
      if (eat('(')) {
        std::vector<double> args;
        if (!eat(')')) {
          do {
            args.push_back(sum());
          } while (eat(','));
          if (!eat(')')) {
            throw BoutException("Missing ')' after arguments of {} in '{}'", id, src);
          }
        }
        return call(id, args);
      }
      return lookup(id);
    }
    throw BoutException("Unexpected character '{}' in expression '{}'", std::string(1, c), src);
  }
};

// Fills of whole fields from an expression (initial profiles, field-valued options) run the x planes on the host's
// cores: fn(x) for x in [0, nx), exceptions rethrown on the calling thread.
void forEachPlane(int nx, const std::function<void(int)>& fn) {
  const unsigned hw = std::thread::hardware_concurrency();
  const int nt = std::max(1, std::min<int>(nx, hw == 0 ? 1 : static_cast<int>(hw)));
  if (nt == 1) {
    for (int x = 0; x < nx; ++x) {
      fn(x);
    }
    return;
  }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> errs(nt);
  for (int t = 0; t < nt; ++t) {
    pool.emplace_back([&, t]() {
      try {
        for (int x = t; x < nx; x += nt) {
          fn(x);
        }
      } catch (...) {
        errs[t] = std::current_exception();
      }
    });
  }
  for (auto& th : pool) {
    th.join();
  }
  for (auto& e : errs) {
    if (e) {
      std::rethrow_exception(e);
    }
  }
}

double evalOption(Options& opt, double x = 0, double y = 0, double z = 0) {
  EvalContext c;
  c.x = x;
  c.y = y;
  c.z = z;
  c.section = opt.parent();
  return Expr(opt.text(), c).eval();
}

} // namespace b200
} // namespace bout

using bout::b200::check;

// ======================================================================================================
// Options
// ======================================================================================================
Options& Options::root() {
  static Options* r = new Options();
  return *r;
}

void Options::cleanup() {
  Options& r = root();
  r.children.clear();
  r.has_value = false;
  r.value.clear();
}

Options& Options::operator[](const std::string& key_in) {
  const auto colon = key_in.find(':');
  if (colon != std::string::npos) {
    return (*this)[key_in.substr(0, colon)][key_in.substr(colon + 1)];
  }
  const std::string& key = key_in; // names are case-sensitive (reference src/sys/options.cxx:109-143)
  auto it = children.find(key);
  if (it == children.end()) {
    it = children.emplace(key, Options()).first;
    it->second.parent_ = this;
    it->second.full_name = full_name.empty() ? key : full_name + ":" + key;
  }
  return it->second;
}

bool Options::isSet(const std::string& key_in) const {
  const auto it = children.find(key_in);
  return it != children.end() && it->second.has_value;
}

template <>
void Options::assignValue<std::string>(const std::string& v) { setText(v); }
template <>
void Options::assignValue<double>(const double& v) { char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", v);
  setText(buf); }
template <>
void Options::assignValue<int>(const int& v) { setText(std::to_string(v)); }
template <>
void Options::assignValue<bool>(const bool& v) { setText(v ? "true" : "false"); }
template <>
void Options::assignValue<Field3D>(const Field3D& v) { field3d = &v; }

// A field-valued option is an expression of x, y, z evaluated over the mesh (reference Options::as<Field3D>,
// src/sys/options.cxx:563-610 -> FieldFactory::create3D; coordinates as in initialProfile below)
template <>
Field3D Options::as<Field3D>() const {
  Mesh* m = bout::globals::mesh;
  if (m == nullptr) {
    throw BoutException("Option {}: a Field3D value needs a mesh", full_name);
  }
  Field3D f(m);
  std::vector<double> h(static_cast<size_t>(m->LocalNx) * m->LocalNy * m->LocalNz);
  bout::b200::forEachPlane(m->LocalNx, [&](int x) {
    for (int y = 0; y < m->LocalNy; ++y) {
      for (int z = 0; z < m->LocalNz; ++z) {
        h[(static_cast<size_t>(x) * m->LocalNy + y) * m->LocalNz + z] = bout::b200::evalOption(
            const_cast<Options&>(*this), m->GlobalX(x), TWOPI * m->GlobalY(y), TWOPI * m->GlobalZ(z));
      }
    }
  });
  f.setHost(h);
  return f;
}

template <>
std::string Options::as<std::string>() const { return value; }
template <>
double Options::as<double>() const { return bout::b200::evalOption(const_cast<Options&>(*this)); }
template <>
int Options::as<int>() const {
  const double v = bout::b200::evalOption(const_cast<Options&>(*this));
  const int r = static_cast<int>(std::lround(v));
  if (std::fabs(v - r) > 1e-3) {
    throw BoutException("Value for option {} = {} is not an integer", full_name, v);
  }
  return r;
}
template <>
bool Options::as<bool>() const {
  std::string s = value;
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); });
  if (s == "true" || s == "yes" || s == "y" || s == "t" || s == "1") {
    return true;
  }
  if (s == "false" || s == "no" || s == "n" || s == "f" || s == "0") {
    return false;
  }
  throw BoutException("Value for option {} = '{}' cannot be converted to a bool", full_name, value);
}

Options& Options::assignRepeat(const Field3D& f, const std::string&, bool, const std::string&) {
  field3d = &f;
  return *this;
}
Options& Options::assignRepeat(const Field2D&, const std::string&, bool, const std::string&) { return *this; }
Options& Options::setAttributes(std::initializer_list<std::pair<std::string, AttributeType>> attrs) {
  for (const auto& a : attrs) {
    attributes[a.first] = a.second;
  }
  return *this;
}

void readOptionsText(const std::string& ini_text, Options& root) {
  std::istringstream in(ini_text);
  std::string line;
  Options* section = &root;
  auto trim = [](std::string s) {
    const auto b = s.find_first_not_of(" \t\r\n");
    if (b == std::string::npos) {
      return std::string();
    }
    const auto e = s.find_last_not_of(" \t\r\n");
    return s.substr(b, e - b + 1);
  };
  auto nextLine = [&](std::string& out) { // stripped of comments and surrounding blanks (options_ini.cxx:179-187)
    if (!std::getline(in, out)) {
      return false;
    }
    const auto hash = out.find_first_of("#;");
    if (hash != std::string::npos) {
      out.erase(hash);
    }
    out = trim(out);
    return true;
  };
  auto openBrackets = [](const std::string& t) { // options_ini.cxx:121-132
    int n = 0;
    for (char ch : t) {
      n += (ch == '(' || ch == '[') ? 1 : ((ch == ')' || ch == ']') ? -1 : 0);
    }
    return n;
  };
  while (nextLine(line)) {
    if (line.empty()) {
      continue;
    }
    if (line.front() == '[') {
      const auto close = line.find(']');
      if (close == std::string::npos) {
        throw BoutException("Options: missing ']' in line '{}'", line);
      }
      section = &root[trim(line.substr(1, close - 1))];
      continue;
    }
    const auto eq = line.find('=');
    if (eq == std::string::npos) { // a bare key is a flag set to true (options_ini.cxx:194-201)
      (*section)[line].setText("true");
      continue;
    }
    std::string val = trim(line.substr(eq + 1));
    for (int open = openBrackets(val); open > 0;) { // a value continues over lines until its brackets balance (:134-150)
      std::string more;
      if (!nextLine(more)) {
        throw BoutException("Options: unbalanced brackets in the value starting '{}'", val);
      }
      open += openBrackets(more);
      val += more;
    }
    if (val.size() >= 2 && val.front() == '"' && val.back() == '"') {
      val = val.substr(1, val.size() - 2);
    }
    (*section)[trim(line.substr(0, eq))].setText(val);
  }
}

// ======================================================================================================
// Fields
// ======================================================================================================
Mesh* FieldBase::getMesh() const { return fieldmesh != nullptr ? fieldmesh : bout::globals::mesh; }

void FieldBase::makeUnique(size_t count) {
  if (data) {
    data->launchDependents();
  }
  if (!data) {
    data = bout::b200::acquire(count);
  } else if (data.use_count() > 1) {
    data->ensure();       // a combination that was never launched is the contents being copied
    data->syncToDevice(); // pending host-side element writes belong to the contents being copied
    auto fresh = bout::b200::acquire(count);
    check(bout_b200_memcpy_d2d(fresh->ptr, data->ptr, count * sizeof(double), nullptr));
    bout::b200::countCowCopy();
    data = fresh;
  }
  data->ensure(); // a partial write lands on the values the block stands for
  // every caller is about to write: the contents get a new identity and lose their recipe
  data->version = bout::b200::nextVersion();
  data->recipe = bout::b200::Block::Recipe{};
}

namespace {
// Host mirror of a field's block for element access: writable = the caller may store through the pointer
double* mirrorOf(const std::shared_ptr<bout::b200::Block>& data, bool writable, const char* what) {
  if (!data) {
    throw BoutException("{}: field has no data", what);
  }
  data->ensure();
  data->syncToHost();
  if (writable) {
    data->mirror = bout::b200::Block::host_newer;
  }
  return data->host.data();
}
} // namespace

namespace {
bout_b200_grid gridOf(const FieldBase& f) { return f.getMesh()->grid(); }

bout_b200_ewterm term(double coef, const FieldBase* f1 = nullptr, bool f1_2d = false, const FieldBase* f2 = nullptr,
                      bool f2_2d = false, const FieldBase* dv = nullptr, bool dv_2d = false) {
  bout_b200_ewterm t{};
  t.coef = coef;
  t.f1 = f1 ? f1->dptr() : nullptr;
  t.f2 = f2 ? f2->dptr() : nullptr;
  t.div = dv ? dv->dptr() : nullptr;
  t.f1_is2d = f1_2d;
  t.f2_is2d = f2_2d;
  t.div_is2d = dv_2d;
  return t;
}

void requireData(const FieldBase& f, const char* what) {
  if (!f.isAllocated()) {
    throw BoutException("{}: field is not allocated", what);
  }
}

// out (a fresh or in-place Field3D) = sum of terms over RGN_ALL
void ew3(Field3D& out, std::initializer_list<bout_b200_ewterm> terms) {
  const auto g = gridOf(out);
  std::vector<bout_b200_ewterm> t(terms);
  check(bout_b200_elementwise(&g, BOUT_B200_RGN_ALL, static_cast<int>(t.size()), t.data(), out.wptr(), nullptr));
}

// Field2D algebra runs through the same kernel on a grid with nz = 1
void ew2(Field2D& out, std::initializer_list<bout_b200_ewterm> terms) {
  Mesh* m = out.getMesh();
  bout_b200_grid g{m->LocalNx, m->LocalNy, 1, m->xstart, m->xend, m->ystart, m->yend};
  std::vector<bout_b200_ewterm> t(terms);
  for (auto& e : t) { // every operand is a "3D" array with nz = 1 here
    e.f1_is2d = e.f2_is2d = e.div_is2d = 0;
  }
  check(bout_b200_elementwise(&g, BOUT_B200_RGN_ALL, static_cast<int>(t.size()), t.data(), out.wptr(), nullptr));
}
} // namespace

Field2D::Field2D(Mesh* m) { fieldmesh = m; }
Field2D::Field2D(BoutReal v, Mesh* m) {
  fieldmesh = m;
  *this = v;
}
size_t Field2D::size() const {
  Mesh* m = getMesh();
  return static_cast<size_t>(m->LocalNx) * m->LocalNy;
}
Field2D& Field2D::allocate() {
  makeUnique(size());
  return *this;
}
double* Field2D::wptr() {
  makeUnique(size());
  data->syncToDevice();
  data->mirror = bout::b200::Block::none; // the device copy is about to change
  return data->ptr;
}
BoutReal& Field2D::operator()(int x, int y) {
  makeUnique(size());
  return mirrorOf(data, true, "Field2D(x,y)")[static_cast<size_t>(x) * getMesh()->LocalNy + y];
}
const BoutReal& Field2D::operator()(int x, int y) const {
  return mirrorOf(data, false, "Field2D(x,y)")[static_cast<size_t>(x) * getMesh()->LocalNy + y];
}
BoutReal& Field2D::operator[](const Ind2D& i) {
  makeUnique(size());
  return mirrorOf(data, true, "Field2D[i]")[i.ind];
}
const BoutReal& Field2D::operator[](const Ind2D& i) const { return mirrorOf(data, false, "Field2D[i]")[i.ind]; }
Field2D& Field2D::operator=(BoutReal v) {
  data = bout::b200::acquire(size());
  Field2D& self = *this;
  ew2(self, {term(v)});
  return *this;
}
void Field2D::setHost(const std::vector<double>& v) {
  if (v.size() != size()) {
    throw BoutException("Field2D::setHost: expected {} values, got {}", size(), v.size());
  }
  data = bout::b200::acquire(size());
  check(bout_b200_memcpy_h2d(data->ptr, v.data(), v.size() * sizeof(double), nullptr));
}
std::vector<double> Field2D::getHost() const {
  requireData(*this, "Field2D::getHost");
  std::vector<double> v(size());
  check(bout_b200_memcpy_d2h(v.data(), dptr(), v.size() * sizeof(double), nullptr));
  return v;
}
Field2D& Field2D::operator+=(const Field2D& r) { return *this = *this + r; }
Field2D& Field2D::operator-=(const Field2D& r) { return *this = *this - r; }
Field2D& Field2D::operator*=(BoutReal s) { return *this = *this * s; }

#define B200_F2_BINOP(OP, ...)                         \
  Field2D operator OP(const Field2D& a, const Field2D& b) { \
    requireData(a, "Field2D operator");                \
    requireData(b, "Field2D operator");                \
    Field2D r(a.getMesh());                            \
    r.allocate();                                      \
    ew2(r, {__VA_ARGS__});                             \
    return r;                                          \
  }
B200_F2_BINOP(+, term(1.0, &a), term(1.0, &b))
B200_F2_BINOP(-, term(1.0, &a), term(-1.0, &b))
B200_F2_BINOP(*, term(1.0, &a, false, &b))
B200_F2_BINOP(/, term(1.0, &a, false, nullptr, false, &b))
Field2D operator*(const Field2D& a, BoutReal b) {
  Field2D r(a.getMesh());
  r.allocate();
  ew2(r, {term(b, &a)});
  return r;
}
Field2D operator*(BoutReal a, const Field2D& b) { return b * a; }
Field2D operator/(const Field2D& a, BoutReal b) { return a * (1.0 / b); }
Field2D operator/(BoutReal a, const Field2D& b) {
  Field2D r(b.getMesh());
  r.allocate();
  ew2(r, {term(a, nullptr, false, nullptr, false, &b)});
  return r;
}
Field2D operator+(const Field2D& a, BoutReal b) {
  Field2D r(a.getMesh());
  r.allocate();
  ew2(r, {term(1.0, &a), term(b)});
  return r;
}
Field2D operator-(const Field2D& a) { return a * (-1.0); }

Field3D::Field3D(Mesh* m) { fieldmesh = m; }
Field3D::Field3D(const Field3D& f) : FieldBase(f) { // shares the data block (field3d.cxx:74-84)
  deriv = nullptr;
}
Field3D::Field3D(BoutReal v, Mesh* m) {
  fieldmesh = m;
  *this = v;
}
Field3D::~Field3D() { delete deriv; }
size_t Field3D::size() const {
  Mesh* m = getMesh();
  return static_cast<size_t>(m->LocalNx) * m->LocalNy * m->LocalNz;
}
Field3D& Field3D::allocate() {
  makeUnique(size());
  return *this;
}
double* Field3D::wptr() {
  makeUnique(size());
  data->syncToDevice();
  data->mirror = bout::b200::Block::none; // the device copy is about to change
  return data->ptr;
}
BoutReal& Field3D::operator()(int x, int y, int z) {
  makeUnique(size());
  Mesh* m = getMesh();
  return mirrorOf(data, true, "Field3D(x,y,z)")[(static_cast<size_t>(x) * m->LocalNy + y) * m->LocalNz + z];
}
const BoutReal& Field3D::operator()(int x, int y, int z) const {
  Mesh* m = getMesh();
  return mirrorOf(data, false, "Field3D(x,y,z)")[(static_cast<size_t>(x) * m->LocalNy + y) * m->LocalNz + z];
}
BoutReal& Field3D::operator[](const Ind3D& i) {
  makeUnique(size());
  return mirrorOf(data, true, "Field3D[i]")[i.ind];
}
const BoutReal& Field3D::operator[](const Ind3D& i) const { return mirrorOf(data, false, "Field3D[i]")[i.ind]; }
const Region<Ind3D>& Field3D::getRegion(const std::string& name) const { return getMesh()->getRegion3D(name); }
Field3D& Field3D::operator=(const Field3D& f) {
  if (this != &f) {
    data = f.data; // share; copy-on-write happens in wptr()
    if (fieldmesh == nullptr) {
      fieldmesh = f.fieldmesh;
    }
  }
  return *this;
}
Field3D& Field3D::operator=(Field3D&& f) noexcept {
  data = std::move(f.data);
  if (fieldmesh == nullptr) {
    fieldmesh = f.fieldmesh;
  }
  return *this;
}
Field3D& Field3D::operator=(const Field2D& f) {
  requireData(f, "Field3D = Field2D");
  data = bout::b200::acquire(size());
  ew3(*this, {term(1.0, &f, true)});
  return *this;
}
Field3D& Field3D::operator=(BoutReal v) {
  data = bout::b200::acquire(size());
  ew3(*this, {term(v)});
  return *this;
}
void Field3D::setHost(const double* v) {
  data = bout::b200::acquire(size());
  check(bout_b200_memcpy_h2d(data->ptr, v, size() * sizeof(double), nullptr));
}
void Field3D::setHost(const std::vector<double>& v) {
  if (v.size() != size()) {
    throw BoutException("Field3D::setHost: expected {} values, got {}", size(), v.size());
  }
  setHost(v.data());
}
void Field3D::getHost(double* out) const {
  requireData(*this, "Field3D::getHost");
  check(bout_b200_memcpy_d2h(out, dptr(), size() * sizeof(double), nullptr));
}
std::vector<double> Field3D::getHost() const {
  std::vector<double> v(size());
  getHost(v.data());
  return v;
}
Field3D* Field3D::timeDeriv() {
  if (deriv == nullptr) {
    deriv = new Field3D(getMesh());
  }
  return deriv;
}

// ---- Field3D algebra: deferred, fused (bout::b200::Block::pending) ------------------------------------------------
namespace {
using LazyTerm = bout::b200::Block::LazyTerm;
using Terms = std::vector<LazyTerm>;
constexpr size_t kMaxTerms = 8; // bout_b200_elementwise takes 1..8 terms

bool lazyAlgebra() { // BOUT_B200_LAZY=0: every operator launches its own kernel as it returns
  static const bool on = [] {
    const char* e = getenv("BOUT_B200_LAZY");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}
LazyTerm fieldTerm(const FieldBase& f, bool is2d, double coef = 1.0) {
  LazyTerm t;
  t.coef = coef;
  t.f1 = f.shared();
  t.f1_2d = is2d;
  return t;
}
LazyTerm constTerm(double v) {
  LazyTerm t;
  t.coef = v;
  return t;
}
// the combination a Field3D stands for: its own pending terms while they were not launched, else the field itself
Terms exprOf(const Field3D& f, bool inline_pending = true) {
  const auto* b = f.block();
  if (inline_pending && lazyAlgebra() && !b->pending.empty()) {
    return b->pending;
  }
  return {fieldTerm(f, false)};
}
bool pureField(const Terms& t) { // coef * f1 and nothing else
  return t.size() == 1 && t[0].f1 && !t[0].f2 && !t[0].dv;
}
Field3D lazyResult(Mesh* m, Terms&& terms) {
  Field3D r(m);
  r.allocate();
  auto* b = r.block();
  const auto g = m->grid();
  const int gi[7] = {g.nx, g.ny, g.nz, g.xstart, g.xend, g.ystart, g.yend};
  std::copy(gi, gi + 7, b->pgrid);
  b->pending = std::move(terms);
  if (!lazyAlgebra()) {
    b->ensure();
    return r;
  }
  for (const auto& t : b->pending) {
    for (const auto* op : {&t.f1, &t.f2, &t.dv}) {
      if (*op) {
        auto& deps = (*op)->dependents;
        if (deps.size() >= 16) { // forget the ones that are gone or were launched meanwhile
          deps.erase(std::remove_if(deps.begin(), deps.end(),
                                    [](const std::weak_ptr<bout::b200::Block>& w) {
                                      auto d = w.lock();
                                      return !d || d->pending.empty();
                                    }),
                     deps.end());
        }
        deps.push_back(r.shared());
      }
    }
  }
  return r;
}
// a + sign * b with as many of the operands' pending terms absorbed as fit into one launch
Field3D lazySum(const Field3D& a, const Terms& ea_all, Terms eb, double sign, Mesh* m) {
  Terms ea = ea_all;
  if (ea.size() + eb.size() > kMaxTerms) { // keep the longer combination a launch of its own
    if (ea.size() >= eb.size()) {
      ea = exprOf(a, false);
    }
    if (ea.size() + eb.size() > kMaxTerms) {
      return Field3D(); // caller falls back to single terms
    }
  }
  for (auto& t : eb) {
    t.coef *= sign;
    ea.push_back(std::move(t));
  }
  return lazyResult(m, std::move(ea));
}
Field3D addSub(const Field3D& a, const Field3D& b, double sign) {
  requireData(a, "Field3D operator");
  requireData(b, "Field3D operator");
  Field3D r = lazySum(a, exprOf(a), exprOf(b), sign, a.getMesh());
  if (!r.isAllocated()) {
    Terms eb = exprOf(b, false);
    r = lazySum(a, exprOf(a), eb, sign, a.getMesh());
    if (!r.isAllocated()) {
      r = lazySum(a, exprOf(a, false), exprOf(b, false), sign, a.getMesh());
    }
  }
  return r;
}
// every term of `e` times (div = false) or divided by (div = true) the field g (3-D or 2-D): possible while each term
// has the slot free; otherwise e is launched on its own first
Field3D mulDiv(const Field3D& a, const FieldBase& gfield, bool g2d, bool div) {
  requireData(a, "Field3D operator");
  requireData(gfield, "Field3D operator");
  Terms e = exprOf(a);
  auto fits = [&](const Terms& ts) {
    for (const auto& t : ts) {
      if (div ? static_cast<bool>(t.dv) : (t.f1 && t.f2)) {
        return false;
      }
    }
    return true;
  };
  if (!fits(e)) {
    e = exprOf(a, false);
  }
  for (auto& t : e) {
    if (div) {
      t.dv = gfield.shared();
      t.dv_2d = g2d;
    } else if (!t.f1) {
      t.f1 = gfield.shared();
      t.f1_2d = g2d;
    } else {
      t.f2 = gfield.shared();
      t.f2_2d = g2d;
    }
  }
  return lazyResult(a.getMesh(), std::move(e));
}
Field3D scaled(const Field3D& f, double s) {
  requireData(f, "Field3D operator");
  Terms e = exprOf(f);
  for (auto& t : e) {
    t.coef *= s;
  }
  return lazyResult(f.getMesh(), std::move(e));
}
Field3D plusConst(const Field3D& f, double fs, double c) { // fs * f + c
  requireData(f, "Field3D operator");
  Terms e = exprOf(f);
  if (e.size() + 1 > kMaxTerms) {
    e = exprOf(f, false);
  }
  for (auto& t : e) {
    t.coef *= fs;
  }
  e.push_back(constTerm(c));
  return lazyResult(f.getMesh(), std::move(e));
}
} // namespace

Field3D operator+(const Field3D& a, const Field3D& b) { return addSub(a, b, 1.0); }
Field3D operator-(const Field3D& a, const Field3D& b) {
  // remembered for the single-index loop dispatcher: r == a - b while a and b keep these versions
  const bool plain = a.isAllocated() && b.isAllocated() && a.block()->pending.empty() && b.block()->pending.empty();
  Field3D r = addSub(a, b, -1.0);
  if (plain) {
    bout::b200::Block::Recipe rec;
    rec.op = 1;
    rec.a = a.block();
    rec.b = b.block();
    rec.a_ver = a.block()->version;
    rec.b_ver = b.block()->version;
    r.block()->recipe = rec;
  }
  return r;
}
Field3D operator*(const Field3D& a, const Field3D& b) {
  requireData(a, "Field3D operator");
  requireData(b, "Field3D operator");
  // the operand that is a plain field multiplies the other one's terms
  if (!b.block()->pending.empty() && a.block()->pending.empty()) {
    return mulDiv(b, a, false, false);
  }
  return mulDiv(a, b, false, false); // (b is launched first when it is a pending combination itself: dptr in ensure)
}
Field3D operator/(const Field3D& a, const Field3D& b) { return mulDiv(a, b, false, true); }
Field3D operator+(const Field3D& a, const Field2D& b) {
  requireData(b, "Field3D operator");
  Field3D r = lazySum(a, exprOf(a), {fieldTerm(b, true)}, 1.0, a.getMesh());
  return r.isAllocated() ? r : lazySum(a, exprOf(a, false), {fieldTerm(b, true)}, 1.0, a.getMesh());
}
Field3D operator-(const Field3D& a, const Field2D& b) {
  requireData(b, "Field3D operator");
  Field3D r = lazySum(a, exprOf(a), {fieldTerm(b, true)}, -1.0, a.getMesh());
  return r.isAllocated() ? r : lazySum(a, exprOf(a, false), {fieldTerm(b, true)}, -1.0, a.getMesh());
}
Field3D operator*(const Field3D& a, const Field2D& b) { return mulDiv(a, b, true, false); }
Field3D operator/(const Field3D& a, const Field2D& b) { return mulDiv(a, b, true, true); }
Field3D operator+(const Field2D& a, const Field3D& b) { return b + a; }
Field3D operator*(const Field2D& a, const Field3D& b) { return b * a; }
Field3D operator-(const Field2D& a, const Field3D& b) {
  requireData(a, "Field3D operator");
  requireData(b, "Field3D operator");
  Terms e = exprOf(b);
  if (e.size() + 1 > kMaxTerms) {
    e = exprOf(b, false);
  }
  for (auto& t : e) {
    t.coef = -t.coef;
  }
  e.insert(e.begin(), fieldTerm(a, true));
  return lazyResult(b.getMesh(), std::move(e));
}
Field3D operator/(const Field2D& a, const Field3D& b) {
  requireData(a, "Field3D operator");
  requireData(b, "Field3D operator");
  LazyTerm t = fieldTerm(a, true);
  t.dv = b.shared(); // (a pending b is launched first)
  return lazyResult(b.getMesh(), {t});
}
Field3D operator+(const Field3D& f, BoutReal s) { return plusConst(f, 1.0, s); }
Field3D operator-(const Field3D& f, BoutReal s) { return plusConst(f, 1.0, -s); }
Field3D operator*(const Field3D& f, BoutReal s) { return scaled(f, s); }
Field3D operator/(const Field3D& f, BoutReal s) { return scaled(f, 1.0 / s); } // generated_fieldops.cxx:543-558
Field3D operator+(BoutReal s, const Field3D& f) { return plusConst(f, 1.0, s); }
Field3D operator-(BoutReal s, const Field3D& f) { return plusConst(f, -1.0, s); }
Field3D operator*(BoutReal s, const Field3D& f) { return scaled(f, s); }
Field3D operator/(BoutReal s, const Field3D& f) {
  requireData(f, "Field3D operator");
  LazyTerm t = constTerm(s);
  t.dv = f.shared();
  return lazyResult(f.getMesh(), {t});
}
Field3D operator-(const Field3D& f) { return scaled(f, -1.0); } // field3d.cxx:581

namespace {
Field3D whereImpl(const FieldBase& test, bool test2d, const Field3D* gt0, double gv, const Field3D* le0, double lv, Mesh* m) {
  requireData(test, "where");
  Field3D r(m);
  const auto g = m->grid();
  const double* a = gt0 != nullptr ? gt0->dptr() : nullptr;
  const double* b = le0 != nullptr ? le0->dptr() : nullptr;
  check(bout_b200_where(&g, test.dptr(), test2d ? 1 : 0, a, gv, b, lv, r.wptr(), nullptr));
  return r;
}
} // namespace
Field3D where(const Field3D& t, const Field3D& a, const Field3D& b) { return whereImpl(t, false, &a, 0.0, &b, 0.0, t.getMesh()); }
Field3D where(const Field3D& t, const Field3D& a, BoutReal b) { return whereImpl(t, false, &a, 0.0, nullptr, b, t.getMesh()); }
Field3D where(const Field3D& t, BoutReal a, const Field3D& b) { return whereImpl(t, false, nullptr, a, &b, 0.0, t.getMesh()); }
Field3D where(const Field3D& t, BoutReal a, BoutReal b) { return whereImpl(t, false, nullptr, a, nullptr, b, t.getMesh()); }
Field3D where(const Field2D& t, const Field3D& a, const Field3D& b) { return whereImpl(t, true, &a, 0.0, &b, 0.0, a.getMesh()); }
Field3D where(const Field2D& t, const Field3D& a, BoutReal b) { return whereImpl(t, true, &a, 0.0, nullptr, b, a.getMesh()); }
Field3D where(const Field2D& t, BoutReal a, const Field3D& b) { return whereImpl(t, true, nullptr, a, &b, 0.0, b.getMesh()); }

Field3D& Field3D::operator+=(const Field3D& r) { return *this = *this + r; }
Field3D& Field3D::operator+=(const Field2D& r) { return *this = *this + r; }
Field3D& Field3D::operator+=(BoutReal r) { return *this = *this + r; }
Field3D& Field3D::operator-=(const Field3D& r) { return *this = *this - r; }
Field3D& Field3D::operator-=(const Field2D& r) { return *this = *this - r; }
Field3D& Field3D::operator-=(BoutReal r) { return *this = *this - r; }
Field3D& Field3D::operator*=(const Field3D& r) { return *this = *this * r; }
Field3D& Field3D::operator*=(const Field2D& r) { return *this = *this * r; }
Field3D& Field3D::operator*=(BoutReal r) { return *this = *this * r; }
Field3D& Field3D::operator/=(const Field3D& r) { return *this = *this / r; }
Field3D& Field3D::operator/=(const Field2D& r) { return *this = *this / r; }
Field3D& Field3D::operator/=(BoutReal r) { return *this = *this / r; }

// ======================================================================================================
// FieldPerp (reference include/bout/fieldperp.hxx, src/field/fieldperp.cxx, generated_fieldops.cxx)
// ======================================================================================================
namespace {
bout_b200_grid perpGridOf(const FieldBase& f) {
  Mesh* m = f.getMesh();
  return bout_b200_grid{m->LocalNx, 1, m->LocalNz, m->xstart, m->xend, 0, 0};
}
// out = sum of terms over the whole plane; every operand is an [nx][nz] array (a "3D" field with ny = 1)
void ewp(FieldPerp& out, std::initializer_list<bout_b200_ewterm> terms) {
  const auto g = perpGridOf(out);
  std::vector<bout_b200_ewterm> t(terms);
  check(bout_b200_elementwise(&g, BOUT_B200_RGN_ALL, static_cast<int>(t.size()), t.data(), out.wptr(), nullptr));
}
FieldPerp perpLike(const FieldPerp& a) {
  FieldPerp r(a.getMesh(), CELL_CENTRE, a.getIndex());
  r.allocate();
  return r;
}
} // namespace

FieldPerp::FieldPerp(Mesh* m, CELL_LOC, int yindex_in) : yindex(yindex_in) { fieldmesh = m; }
FieldPerp::FieldPerp(BoutReal v, Mesh* m) {
  fieldmesh = m;
  *this = v;
}
int FieldPerp::getNx() const { return getMesh()->LocalNx; }
int FieldPerp::getNz() const { return getMesh()->LocalNz; }
size_t FieldPerp::size() const { return static_cast<size_t>(getNx()) * getNz(); }
FieldPerp& FieldPerp::allocate() {
  makeUnique(size());
  return *this;
}
double* FieldPerp::wptr() {
  makeUnique(size());
  data->syncToDevice();
  data->mirror = bout::b200::Block::none; // the device copy is about to change
  return data->ptr;
}
FieldPerp& FieldPerp::operator=(BoutReal v) {
  data = bout::b200::acquire(size());
  ewp(*this, {term(v)});
  return *this;
}
void FieldPerp::setHost(const double* v) {
  data = bout::b200::acquire(size());
  check(bout_b200_memcpy_h2d(data->ptr, v, size() * sizeof(double), nullptr));
}
std::vector<double> FieldPerp::getHost() const {
  requireData(*this, "FieldPerp::getHost");
  std::vector<double> v(size());
  check(bout_b200_memcpy_d2h(v.data(), dptr(), size() * sizeof(double), nullptr));
  return v;
}
BoutReal& FieldPerp::operator()(int x, int z) {
  makeUnique(size());
  return mirrorOf(data, true, "FieldPerp(x,z)")[static_cast<size_t>(x) * getNz() + z];
}
const BoutReal& FieldPerp::operator()(int x, int z) const {
  return mirrorOf(data, false, "FieldPerp(x,z)")[static_cast<size_t>(x) * getNz() + z];
}

FieldPerp sliceXZ(const Field3D& f, int y) {
  requireData(f, "sliceXZ");
  FieldPerp r(f.getMesh(), CELL_CENTRE, y);
  const auto g = gridOf(f);
  check(bout_b200_copy_xz(&g, y, const_cast<double*>(f.dptr()), r.wptr(), 1, nullptr));
  return r;
}
Field3D& Field3D::operator=(const FieldPerp& p) {
  requireData(p, "Field3D = FieldPerp");
  const auto g = gridOf(*this);
  check(bout_b200_copy_xz(&g, p.getIndex(), wptr(), const_cast<double*>(p.dptr()), 0, nullptr));
  return *this;
}

#define B200_FP_BINOP(OP, ...)                                  \
  FieldPerp operator OP(const FieldPerp& a, const FieldPerp& b) { \
    requireData(a, "FieldPerp operator");                       \
    requireData(b, "FieldPerp operator");                       \
    FieldPerp r = perpLike(a);                                  \
    ewp(r, {__VA_ARGS__});                                      \
    return r;                                                   \
  }
B200_FP_BINOP(+, term(1.0, &a), term(1.0, &b))
B200_FP_BINOP(-, term(1.0, &a), term(-1.0, &b))
B200_FP_BINOP(*, term(1.0, &a, false, &b))
B200_FP_BINOP(/, term(1.0, &a, false, nullptr, false, &b))
#undef B200_FP_BINOP
FieldPerp operator+(const FieldPerp& a, const Field3D& b) { return a + sliceXZ(b, a.getIndex()); }
FieldPerp operator-(const FieldPerp& a, const Field3D& b) { return a - sliceXZ(b, a.getIndex()); }
FieldPerp operator*(const FieldPerp& a, const Field3D& b) { return a * sliceXZ(b, a.getIndex()); }
FieldPerp operator/(const FieldPerp& a, const Field3D& b) { return a / sliceXZ(b, a.getIndex()); }
#define B200_FP_SCALAR(...)                 \
  {                                         \
    requireData(f, "FieldPerp operator");   \
    FieldPerp r = perpLike(f);              \
    ewp(r, {__VA_ARGS__});                  \
    return r;                               \
  }
FieldPerp operator+(const FieldPerp& f, BoutReal s) B200_FP_SCALAR(term(1.0, &f), term(s))
FieldPerp operator-(const FieldPerp& f, BoutReal s) B200_FP_SCALAR(term(1.0, &f), term(-s))
FieldPerp operator*(const FieldPerp& f, BoutReal s) B200_FP_SCALAR(term(s, &f))
FieldPerp operator/(const FieldPerp& f, BoutReal s) B200_FP_SCALAR(term(1.0 / s, &f)) // generated_fieldops.cxx: * (1 / rhs)
FieldPerp operator+(BoutReal s, const FieldPerp& f) B200_FP_SCALAR(term(s), term(1.0, &f))
FieldPerp operator-(BoutReal s, const FieldPerp& f) B200_FP_SCALAR(term(s), term(-1.0, &f))
FieldPerp operator*(BoutReal s, const FieldPerp& f) B200_FP_SCALAR(term(s, &f))
FieldPerp operator/(BoutReal s, const FieldPerp& f) B200_FP_SCALAR(term(s, nullptr, false, nullptr, false, &f))
FieldPerp operator-(const FieldPerp& f) B200_FP_SCALAR(term(-1.0, &f))
#undef B200_FP_SCALAR
FieldPerp& FieldPerp::operator+=(const FieldPerp& r) { return *this = *this + r; }
FieldPerp& FieldPerp::operator-=(const FieldPerp& r) { return *this = *this - r; }
FieldPerp& FieldPerp::operator*=(const FieldPerp& r) { return *this = *this * r; }
FieldPerp& FieldPerp::operator/=(const FieldPerp& r) { return *this = *this / r; }
FieldPerp& FieldPerp::operator+=(BoutReal r) { return *this = *this + r; }
FieldPerp& FieldPerp::operator-=(BoutReal r) { return *this = *this - r; }
FieldPerp& FieldPerp::operator*=(BoutReal r) { return *this = *this * r; }
FieldPerp& FieldPerp::operator/=(BoutReal r) { return *this = *this / r; }
namespace {
FieldPerp unaryPerp(int op, const FieldPerp& f, double arg = 0.0) {
  requireData(f, "FieldPerp function");
  FieldPerp r = perpLike(f);
  const auto g = perpGridOf(f);
  check(bout_b200_unary(&g, op, f.dptr(), arg, r.wptr(), nullptr));
  return r;
}
} // namespace
FieldPerp sqrt(const FieldPerp& f) { return unaryPerp(1, f); }
FieldPerp abs(const FieldPerp& f) { return unaryPerp(2, f); }
FieldPerp exp(const FieldPerp& f) { return unaryPerp(3, f); }
FieldPerp SQ(const FieldPerp& f) { return unaryPerp(8, f); }

namespace {
Field3D unaryOp(int op, const Field3D& f, double arg = 0.0) {
  requireData(f, "Field3D function");
  Field3D r(f.getMesh());
  const auto g = gridOf(f);
  check(bout_b200_unary(&g, op, f.dptr(), arg, r.wptr(), nullptr));
  return r;
}
} // namespace
Field3D sqrt(const Field3D& f) { return unaryOp(1, f); }
Field3D abs(const Field3D& f) { return unaryOp(2, f); }
Field3D exp(const Field3D& f) { return unaryOp(3, f); }
Field3D log(const Field3D& f) { return unaryOp(4, f); }
Field3D sin(const Field3D& f) { return unaryOp(5, f); }
Field3D cos(const Field3D& f) { return unaryOp(6, f); }
Field3D tanh(const Field3D& f) { return unaryOp(7, f); }
Field3D SQ(const Field3D& f) { return unaryOp(8, f); }
Field3D pow(const Field3D& f, BoutReal e) { return unaryOp(9, f, e); }

Field2D DC(const Field3D& f) {
  requireData(f, "DC");
  Field2D r(f.getMesh());
  const auto g = gridOf(f);
  check(bout_b200_dc(&g, f.dptr(), r.wptr(), nullptr));
  return r;
}
Field3D filter(const Field3D& f, int N0) {
  Field3D r(f.getMesh());
  const auto g = gridOf(f);
  check(bout_b200_zfilter(&g, BOUT_B200_RGN_ALL, N0, f.dptr(), r.wptr(), nullptr));
  return r;
}
Field3D lowPass(const Field3D& f, int zmax, bool keep_zonal) {
  Field3D r(f.getMesh());
  const auto g = gridOf(f);
  check(bout_b200_zlowpass(&g, BOUT_B200_RGN_ALL, zmax, keep_zonal ? 1 : 0, f.dptr(), r.wptr(), nullptr));
  return r;
}

// averageY (src/physics/smoothing.cxx:238-323): mean over ystart..yend broadcast to all y.  The y-mean of
// a Field2D is a reduction over a few hundred numbers per x; it runs through the DC kernel by viewing
// the Field2D [nx][ny] as a field with one "y" row per x and z = the interior y range.
Field2D averageY(const Field2D& f) {
  requireData(f, "averageY");
  Mesh* m = f.getMesh();
  const int nyint = m->yend - m->ystart + 1;
  Field2D r(m);
  if (m->NYPE > 1) { // local sums per x, summed over the ranks of the Y column (smoothing.cxx:251-262)
    const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    const size_t nx = static_cast<size_t>(m->LocalNx);
    if (yd.allreduce_y == nullptr || yd.gather_send == nullptr || yd.capacity < nx) {
      throw BoutException("averageY on NYPE = {} ranks needs YDecomposition::allreduce_y and gather_send", yd.nype);
    }
    std::vector<double> h = f.getHost(), out(h.size()), s(nx, 0.0);
    for (size_t x = 0; x < nx; ++x) {
      for (int y = m->ystart; y <= m->yend; ++y) {
        s[x] += h[x * m->LocalNy + y];
      }
    }
    check(bout_b200_memcpy_h2d(yd.gather_send, s.data(), nx * sizeof(double), nullptr));
    if (yd.allreduce_y(nx, 0, yd.user) != 0) {
      throw BoutException("averageY: the all-reduce over the Y column failed");
    }
    check(bout_b200_memcpy_d2h(s.data(), yd.gather_send, nx * sizeof(double), nullptr));
    check(bout_b200_stream_sync(nullptr));
    for (size_t x = 0; x < nx; ++x) {
      for (int y = 0; y < m->LocalNy; ++y) {
        out[x * m->LocalNy + y] = s[x] / (nyint * m->NYPE);
      }
    }
    r.setHost(out);
    return r;
  }
  if (nyint == m->LocalNy) { // no y guards: every row is one contiguous run
    bout_b200_grid g{m->LocalNx, 1, m->LocalNy, 0, m->LocalNx - 1, 0, 0};
    Field2D mean(m); // holds nx values in its first nx slots
    check(bout_b200_dc(&g, f.dptr(), mean.wptr(), nullptr));
    // broadcast: r[x][y] = mean[x] -> elementwise with f1 a "Field2D" on the (nx, 1, ny) grid
    bout_b200_ewterm t{};
    t.coef = 1.0;
    t.f1 = mean.dptr();
    t.f1_is2d = 1;
    check(bout_b200_elementwise(&g, BOUT_B200_RGN_ALL, 1, &t, r.wptr(), nullptr));
    return r;
  }
  // with y guards: gather the interior rows through the host-free path of a strided view is not
  // available in the C-ABI; this is a set-up/diagnostic-sized array, reduce it on the host
  std::vector<double> h = f.getHost(), out(h.size());
  for (int x = 0; x < m->LocalNx; ++x) {
    double s = 0.0;
    for (int y = m->ystart; y <= m->yend; ++y) {
      s += h[static_cast<size_t>(x) * m->LocalNy + y];
    }
    s /= nyint;
    for (int y = 0; y < m->LocalNy; ++y) {
      out[static_cast<size_t>(x) * m->LocalNy + y] = s;
    }
  }
  r.setHost(out);
  return r;
}

// ---- boundaries ---------------------------------------------------------------------------------------
namespace {
// The bout_b200_boundary arguments of one condition on one side of one field; false = nothing to apply.
bool boundarySpec(Field3D& f, const std::string& cond_in, int side, bout_b200_bndry_spec& spec) {
  std::string cond = cond_in;
  cond.erase(std::remove_if(cond.begin(), cond.end(), [](unsigned char c) { return std::isspace(c); }), cond.end());
  std::transform(cond.begin(), cond.end(), cond.begin(), [](unsigned char c) { return std::tolower(c); });
  if (cond.empty() || cond == "none" || cond == "null") {
    return false;
  }
  if (cond.rfind("relax(", 0) == 0 && cond.back() == ')') {
    // relax(op[, rate]) (BoundaryRelax, boundary_standard.cxx:3599-3637): apply() is the wrapped condition's; the
    // relaxation only enters apply_ddt, which the solver calls for evolve_bndry = true (not supported here)
    const std::string inner = cond.substr(6, cond.size() - 7);
    int depth = 0;
    size_t cut = inner.size();
    for (size_t i = 0; i < inner.size(); ++i) {
      depth += inner[i] == '(' ? 1 : (inner[i] == ')' ? -1 : 0);
      if (inner[i] == ',' && depth == 0) {
        cut = i;
        break;
      }
    }
    return boundarySpec(f, inner.substr(0, cut), side, spec);
  }
  double value = 0.0;
  const auto paren = cond.find('(');
  if (paren != std::string::npos) {
    const std::string arg = cond.substr(paren + 1, cond.rfind(')') - paren - 1);
    cond = cond.substr(0, paren);
    if (!arg.empty()) {
      bout::b200::EvalContext c;
      value = bout::b200::Expr(arg, c).eval();
      // the reference evaluates the argument per boundary cell and per time (FieldGenerator); here it must be a constant
      bout::b200::EvalContext c2;
      c2.x = 0.37;
      c2.y = 1.1;
      c2.z = 2.3;
      c2.t = 0.7;
      const double v2 = bout::b200::Expr(arg, c2).eval();
      if (!(v2 == value) && !(std::isnan(v2) && std::isnan(value))) {
        throw BoutException("Boundary condition '{}': a value that depends on x, y, z or t is not available in the B200 "
                            "host layer (constant values only)", cond_in);
      }
    }
  }
  int kind;
  if (cond == "dirichlet" || cond == "dirichlet_o2") {
    kind = 0;
  } else if (cond == "neumann" || cond == "neumann_o2") {
    kind = 1;
  } else if (cond == "free_o2") {
    kind = 4;
  } else if (cond == "free_o3") {
    kind = 5;
  } else {
    throw BoutException("Boundary condition '{}' is not available in the B200 host layer", cond);
  }
  Coordinates* c = f.getMesh()->getCoordinates();
  spec = bout_b200_bndry_spec{};
  spec.f = f.wptr();
  spec.kind = kind;
  spec.side = side;
  spec.value = value;
  spec.spacing2d = side < 2 ? c->dx.dptr() : c->dy.dptr();
  spec.lo = -1;
  spec.hi = -1;
  return true;
}
void applyOne(Field3D& f, const std::string& cond, int side) {
  bout_b200_bndry_spec q;
  if (boundarySpec(f, cond, side, q)) {
    const auto g = f.getMesh()->grid();
    check(bout_b200_boundary(&g, q.kind, q.side, q.value, q.spacing2d, q.lo, q.hi, q.f, nullptr));
  }
}
bool hasBoundary(Mesh* m, int side) {
  if (side < 2) { // only the ranks at the ends of an X row own a physical x boundary (boutmesh.cxx:1019-1059)
    return !m->periodicX && (side == 0 ? m->firstX() : m->lastX());
  }
  return !m->periodicY;
}
} // namespace

void Field3D::setBoundary(const std::string& var) {
  // BoundaryFactory::createFromOptions (src/mesh/boundary_factory.cxx:228-340): per boundary region the keys
  // bndry_<region label>, bndry_<side>, bndry_all are tried in [var], then the same three in [all]; a region for
  // which nothing is set gets "dirichlet" (to prevent undefined guard cells; "none" switches a side off).
  // Region labels of BoutMesh on the default (all-core) topology (boutmesh.cxx:1019-1078): the inner x boundary is
  // "core", the outer one "sol", the y boundaries "lower_target" / "upper_target".
  static const char* labels[4] = {"core", "sol", "lower_target", "upper_target"};
  static const char* sides[4] = {"xin", "xout", "ydown", "yup"};
  bndry_conditions.assign(4, "dirichlet");
  Options& vo = Options::root()[var];
  Options& all = Options::root()["all"];
  for (int s = 0; s < 4; ++s) {
    bool found = false;
    for (Options* o : {&vo, &all}) {
      for (const std::string& key : {std::string("bndry_") + labels[s], std::string("bndry_") + sides[s], std::string("bndry_all")}) {
        if (o->isSet(key)) {
          bndry_conditions[s] = (*o)[key].text();
          found = true;
          break;
        }
      }
      if (found) {
        break;
      }
    }
  }
}
void Field3D::applyBoundary(bool) {
  Mesh* m = getMesh();
  for (int s = 0; s < 4 && s < static_cast<int>(bndry_conditions.size()); ++s) {
    if (hasBoundary(m, s)) {
      applyOne(*this, bndry_conditions[s], s);
    }
  }
}
void Field3D::applyBoundary(const std::string& condition) {
  Mesh* m = getMesh();
  for (int s = 0; s < 4; ++s) {
    if (hasBoundary(m, s)) {
      applyOne(*this, condition, s);
    }
  }
}
void Field3D::applyTDerivBoundary() {
  Mesh* m = getMesh();
  const auto g = m->grid();
  for (int s = 0; s < 4; ++s) {
    if (hasBoundary(m, s) && deriv != nullptr && deriv->isAllocated()) {
      check(bout_b200_boundary(&g, 2, s, 0.0, nullptr, -1, -1, deriv->wptr(), nullptr));
    }
  }
}

Field3D averageY(const Field3D& f) {
  // smoothing.cxx:272-323 on one rank: y-mean per (x,z) broadcast to every y
  requireData(f, "averageY");
  Mesh* m = f.getMesh();
  if (m->LocalNy == 1) {
    return f;
  }
  // mean over ystart..yend per (x, z), then every y row (guards included) set to it: two launches with the place for
  // the all-reduce of a Y communicator between them (bout_b200.h)
  const auto g = m->grid();
  const size_t nxz = static_cast<size_t>(m->LocalNx) * m->LocalNz;
  Field3D r(m);
  if (m->NYPE > 1) {
    const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    if (yd.allreduce_y == nullptr || yd.gather_send == nullptr || yd.capacity < nxz) {
      throw BoutException("averageY on NYPE = {} ranks needs YDecomposition::allreduce_y and gather_send of nx nz doubles",
                          yd.nype);
    }
    check(bout_b200_ymean(&g, f.dptr(), yd.gather_send, nullptr));
    if (yd.allreduce_y(nxz, 0, yd.user) != 0) {
      throw BoutException("averageY: the all-reduce over the Y column failed");
    }
    check(bout_b200_yspread(&g, yd.gather_send, 1.0 / m->NYPE, r.wptr(), nullptr));
    return r;
  }
  auto xz = bout::b200::acquire(nxz);
  check(bout_b200_ymean(&g, f.dptr(), xz->ptr, nullptr));
  check(bout_b200_yspread(&g, xz->ptr, 1.0, r.wptr(), nullptr));
  return r;
}

Field3D smooth_x(const Field3D& f) {
  // smoothing.cxx:46-73: 0.5 f + 0.25 (f(x-1) + f(x+1)) for 0 < x < LocalNx - 1 over all y, z; the two edge planes
  // copied; then the guard cells of the result exchanged
  requireData(f, "smooth_x");
  Mesh* m = f.getMesh();
  const auto g = m->grid();
  Field3D r(m);
  check(bout_b200_smooth_x(&g, f.dptr(), r.wptr(), nullptr));
  m->communicate(r);
  return r;
}

// ======================================================================================================
// Mesh and Coordinates
// ======================================================================================================
const Region<Ind3D>& Mesh::getRegion3D(const std::string& name) { // mesh.cxx:650-665
  auto it = regions3d.find(name);
  if (it != regions3d.end()) {
    return it->second;
  }
  int xs = 0, xe = LocalNx - 1, ys = 0, ye = LocalNy - 1;
  if (name == "RGN_NOBNDRY" || name == "RGN_NOX") {
    xs = xstart;
    xe = xend;
  }
  if (name == "RGN_NOBNDRY" || name == "RGN_NOY") {
    ys = ystart;
    ye = yend;
  }
  if (name != "RGN_ALL" && name != "RGN_NOBNDRY" && name != "RGN_NOX" && name != "RGN_NOY") {
    throw BoutException("Couldn't find region {:s} in regionMap3D", name);
  }
  Region<Ind3D> r;
  for (int x = xs; x <= xe; ++x) {
    for (int y = ys; y <= ye; ++y) {
      for (int z = 0; z < LocalNz; ++z) {
        r.indices.push_back({(x * LocalNy + y) * LocalNz + z, LocalNy, LocalNz});
      }
    }
  }
  return regions3d.emplace(name, std::move(r)).first->second;
}

Mesh* Mesh::create(Options* opts) { return new Mesh(opts != nullptr ? *opts : Options::root()["mesh"]); }

Mesh::Mesh(Options& opts) : options(&opts) {
  Options& root = Options::root();
  const int nx = opts["nx"].withDefault(1), ny = opts["ny"].withDefault(1);
  int nz = 1;
  if (opts.isSet("nz")) {
    nz = opts["nz"].as<int>();
  } else if (root.isSet("MZ")) {
    nz = root["MZ"].as<int>();
  }
  MXG = root.isSet("MXG") ? root["MXG"].as<int>() : opts["MXG"].withDefault(2);
  MYG = root.isSet("MYG") ? root["MYG"].as<int>() : opts["MYG"].withDefault(2);
  if (ny == 1 && !root.isSet("MYG") && !opts.isSet("MYG")) {
    MYG = 0; // boutmesh.cxx:509-517
  }
  periodicX = root.isSet("periodicX") ? root["periodicX"].as<bool>() : opts["periodicX"].withDefault(false);
  symmetricGlobalX = opts["symmetricGlobalX"].withDefault(true);
  symmetricGlobalY = opts["symmetricGlobalY"].withDefault(true);
  const int ixseps1 = opts["ixseps1"].withDefault(nx), ixseps2 = opts["ixseps2"].withDefault(nx);
  periodicY = (ixseps1 >= nx - MXG) && (ixseps2 >= nx - MXG); // closed field lines everywhere
  if (!periodicY && ixseps1 > MXG) {
    throw BoutException("B200 host layer: meshes with a separatrix (ixseps1 = {}) are not supported", ixseps1);
  }
  GlobalNx = nx;
  GlobalNy = ny + 2 * MYG;
  GlobalNz = nz;
  LocalNx = nx;
  LocalNy = ny + 2 * MYG;
  LocalNz = nz;
  {
    // NYPE ranks along y, NXPE = 1 (boutmesh.cxx:255-341): equal slabs of ny / NYPE rows, rank PE_YIND holds global
    // rows PE_YIND * MYSUB ... + MYSUB - 1 plus MYG guard rows either side
    const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    NYPE = yd.nype;
    PE_YIND = yd.yind;
    if (NYPE > 1) {
      if (PE_YIND < 0 || PE_YIND >= NYPE || ny % NYPE != 0) {
        throw BoutException("Mesh: ny = {} does not divide into NYPE = {} slabs (rank {})", ny, NYPE, PE_YIND);
      }
      const int mysub = ny / NYPE;
      if (MYG < 1 || mysub < MYG) {
        throw BoutException("Mesh: NYPE = {} needs MYG >= 1 and at least MYG rows per rank (ny / NYPE = {}, MYG = {})", NYPE,
                            mysub, MYG);
      }
      if (!periodicY) {
        throw BoutException("B200 host layer: NYPE > 1 is implemented for closed field lines (periodic y) only");
      }
      if (yd.exchange == nullptr || yd.send_lo == nullptr || yd.send_hi == nullptr || yd.recv_lo == nullptr
          || yd.recv_hi == nullptr || yd.capacity < static_cast<size_t>(nx) * MYG * nz) {
        throw BoutException("Mesh: NYPE = {} without an exchange function / buffers of at least nx MYG nz doubles", NYPE);
      }
      LocalNy = mysub + 2 * MYG;
      OffsetY = PE_YIND * mysub;
    }
    NXPE = yd.nxpe;
    PE_XIND = yd.xind;
    if (NXPE > 1) {
      const int mx = nx - 2 * MXG;
      if (PE_XIND < 0 || PE_XIND >= NXPE || MXG < 1 || mx % NXPE != 0 || mx / NXPE < MXG) {
        throw BoutException("Mesh: nx - 2 MXG = {} does not divide into NXPE = {} slabs of at least MXG columns (rank {})",
                            mx, NXPE, PE_XIND);
      }
      if (periodicX) {
        throw BoutException("B200 host layer: NXPE > 1 with periodicX is not implemented");
      }
      const size_t plane = static_cast<size_t>(MXG) * LocalNy * nz;
      if (yd.exchange == nullptr || yd.allgather == nullptr || yd.send_lo == nullptr || yd.send_hi == nullptr
          || yd.recv_lo == nullptr || yd.recv_hi == nullptr || yd.gather_send == nullptr || yd.gather_recv == nullptr
          || yd.capacity < plane) {
        throw BoutException("Mesh: NXPE = {} without exchange / all-gather functions and buffers of at least MXG ny nz doubles",
                            NXPE);
      }
      LocalNx = mx / NXPE + 2 * MXG;
      OffsetX = PE_XIND * (mx / NXPE);
    }
  }
  xstart = MXG;
  xend = LocalNx - MXG - 1;
  ystart = MYG;
  yend = LocalNy - MYG - 1;
  zend = nz - 1;
  if (xend < xstart) {
    throw BoutException("Mesh: nx = {} leaves no interior points with MXG = {}", nx, MXG);
  }
  // derivative method defaults (deriv_store.hxx:441-478) overridden by [mesh:ddx] first/second/upwind/flux
  for (const char* dir : {"ddx", "ddy", "ddz"}) {
    methods[std::string(dir) + ":first"] = "C2";
    methods[std::string(dir) + ":second"] = "C2";
    methods[std::string(dir) + ":fourth"] = "C2";
    methods[std::string(dir) + ":upwind"] = "U1";
    methods[std::string(dir) + ":flux"] = "U1";
    for (const char* kind : {"first", "second", "fourth", "upwind", "flux"}) {
      for (Options* sec : {&opts["diff"], &opts[dir]}) { // mesh:diff then mesh:ddx (deriv_store.hxx:308-389)
        if (sec->isSet(kind)) {
          std::string v = (*sec)[kind].text();
          std::transform(v.begin(), v.end(), v.begin(), [](unsigned char c) { return std::toupper(c); });
          methods[std::string(dir) + ":" + kind] = v;
        }
      }
    }
  }
  bout::globals::mesh = this;
  coords = std::make_unique<Coordinates>(this);
}

Mesh::~Mesh() {
  if (bout::globals::mesh == this) {
    bout::globals::mesh = nullptr;
  }
}

std::string Mesh::method(char dir, const std::string& kind) const {
  const std::string key = std::string("dd") + dir + ":" + kind;
  const auto it = methods.find(key);
  return it == methods.end() ? std::string("C2") : it->second;
}

// boutmesh.cxx:3159-3215 on one rank, all-core topology (jyseps1_1 = -1, no separatrix)
BoutReal Mesh::GlobalX(int jx) const {
  const int MX = GlobalNx - 2 * MXG;
  jx += OffsetX; // XGLOBAL (boutmesh.cxx:3159-3190)
  if (!symmetricGlobalX) {
    return static_cast<BoutReal>(jx) / static_cast<BoutReal>(MX);
  }
  return (0.5 + jx - (GlobalNx - MX) * 0.5) / static_cast<BoutReal>(MX);
}
BoutReal Mesh::GlobalY(int jy) const {
  const int ny = GlobalNy - 2 * MYG;
  jy += OffsetY; // YGLOBAL (boutmesh.cxx:3192-3215): the same number on whichever rank holds the row
  if (!symmetricGlobalY) {
    return static_cast<BoutReal>(jy - MYG) / static_cast<BoutReal>(ny);
  }
  return (jy - MYG + 0.5) / static_cast<BoutReal>(ny);
}

namespace bout {
namespace b200 {
YDecomposition& yDecomposition() {
  static YDecomposition yd;
  return yd;
}
} // namespace b200
} // namespace bout

namespace {
// Y guard rows of `n` device fields from the neighbouring ranks: pack -> exchange -> unpack, as many fields per round
// as the exchange buffers hold.  All x (guard columns too): what the single-rank wrap copies (mesh.cxx:351-362).
void exchangeYDevice(const bout_b200_grid& g, int myg, const std::vector<double*>& ptrs) {
  const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
  const size_t per_field = static_cast<size_t>(g.nx) * myg * g.nz;
  const size_t per_round = std::min<size_t>(8, yd.capacity / per_field);
  for (size_t i = 0; i < ptrs.size(); i += per_round) {
    const int n = static_cast<int>(std::min(per_round, ptrs.size() - i));
    check(bout_b200_halo_pack_y(&g, n, ptrs.data() + i, g.ystart, myg, 0, g.nx - 1, yd.send_lo, nullptr));
    check(bout_b200_halo_pack_y(&g, n, ptrs.data() + i, g.yend - myg + 1, myg, 0, g.nx - 1, yd.send_hi, nullptr));
    if (yd.exchange(per_field * n, 0, yd.user) != 0) {
      throw BoutException("Mesh::communicate: the Y exchange between ranks failed (rank {} of {})", yd.yind, yd.nype);
    }
    check(bout_b200_halo_unpack_y(&g, n, ptrs.data() + i, g.ystart - myg, myg, 0, g.nx - 1, yd.recv_lo, nullptr));
    check(bout_b200_halo_unpack_y(&g, n, ptrs.data() + i, g.yend + 1, myg, 0, g.nx - 1, yd.recv_hi, nullptr));
  }
}
} // namespace

namespace {
// X guard planes of device fields from the neighbouring ranks (sendX / wait, boutmesh.cxx:1245-1340): planes are
// contiguous ([x][y][z], x slowest), so "packing" is one device copy per field and side; all y (the y guard rows just
// exchanged or wrapped travel with them: corner cells as in mesh.cxx:351-362).
void exchangeXDevice(Mesh* m, const std::vector<double*>& ptrs) {
  const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
  const size_t plane = static_cast<size_t>(m->MXG) * m->LocalNy * m->LocalNz;
  const size_t per_round = std::max<size_t>(1, yd.capacity / plane);
  const size_t bytes = plane * sizeof(double);
  const size_t row = static_cast<size_t>(m->LocalNy) * m->LocalNz;
  for (size_t i = 0; i < ptrs.size(); i += per_round) {
    const size_t n = std::min(per_round, ptrs.size() - i);
    for (size_t k = 0; k < n; ++k) {
      check(bout_b200_memcpy_d2d(yd.send_lo + k * plane, ptrs[i + k] + m->xstart * row, bytes, nullptr));
      check(bout_b200_memcpy_d2d(yd.send_hi + k * plane, ptrs[i + k] + (m->xend - m->MXG + 1) * row, bytes, nullptr));
    }
    if (yd.exchange(plane * n, 1, yd.user) != 0) {
      throw BoutException("Mesh::communicate: the X exchange between ranks failed (rank {} of {})", yd.xind, yd.nxpe);
    }
    for (size_t k = 0; k < n; ++k) {
      if (!m->firstX()) {
        check(bout_b200_memcpy_d2d(ptrs[i + k], yd.recv_lo + k * plane, bytes, nullptr));
      }
      if (!m->lastX()) {
        check(bout_b200_memcpy_d2d(ptrs[i + k] + (m->xend + 1) * row, yd.recv_hi + k * plane, bytes, nullptr));
      }
    }
  }
}
} // namespace

void Mesh::exchangeX2D(std::vector<double>& f) {
  const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
  const size_t n = static_cast<size_t>(MXG) * LocalNy;
  check(bout_b200_memcpy_h2d(yd.send_lo, f.data() + static_cast<size_t>(xstart) * LocalNy, n * sizeof(double), nullptr));
  check(bout_b200_memcpy_h2d(yd.send_hi, f.data() + static_cast<size_t>(xend - MXG + 1) * LocalNy, n * sizeof(double), nullptr));
  if (yd.exchange(n, 1, yd.user) != 0) {
    throw BoutException("Mesh: the X exchange of a metric array between ranks failed (rank {} of {})", yd.xind, yd.nxpe);
  }
  if (!firstX()) {
    check(bout_b200_memcpy_d2h(f.data(), yd.recv_lo, n * sizeof(double), nullptr));
  }
  if (!lastX()) {
    check(bout_b200_memcpy_d2h(f.data() + static_cast<size_t>(xend + 1) * LocalNy, yd.recv_hi, n * sizeof(double), nullptr));
  }
  check(bout_b200_stream_sync(nullptr));
}

// Metric arrays live on the host while the coordinates are set up: their guard rows take the same route as a field's
// (through the exchange buffers), once per Coordinates::geometry(), not per step.
void Mesh::exchangeY2D(std::vector<double>& f) {
  const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
  const size_t n = static_cast<size_t>(LocalNx) * MYG;
  std::vector<double> lo(n), hi(n);
  for (int x = 0; x < LocalNx; ++x) {
    for (int i = 0; i < MYG; ++i) {
      lo[static_cast<size_t>(x) * MYG + i] = f[static_cast<size_t>(x) * LocalNy + ystart + i];
      hi[static_cast<size_t>(x) * MYG + i] = f[static_cast<size_t>(x) * LocalNy + yend - MYG + 1 + i];
    }
  }
  check(bout_b200_memcpy_h2d(yd.send_lo, lo.data(), n * sizeof(double), nullptr));
  check(bout_b200_memcpy_h2d(yd.send_hi, hi.data(), n * sizeof(double), nullptr));
  if (yd.exchange(n, 0, yd.user) != 0) {
    throw BoutException("Mesh: the Y exchange of a metric array between ranks failed (rank {} of {})", yd.yind, yd.nype);
  }
  check(bout_b200_memcpy_d2h(lo.data(), yd.recv_lo, n * sizeof(double), nullptr));
  check(bout_b200_memcpy_d2h(hi.data(), yd.recv_hi, n * sizeof(double), nullptr));
  check(bout_b200_stream_sync(nullptr));
  for (int x = 0; x < LocalNx; ++x) {
    for (int i = 0; i < MYG; ++i) {
      f[static_cast<size_t>(x) * LocalNy + i] = lo[static_cast<size_t>(x) * MYG + i];
      f[static_cast<size_t>(x) * LocalNy + yend + 1 + i] = hi[static_cast<size_t>(x) * MYG + i];
    }
  }
}
BoutReal Mesh::GlobalZ(int jz) const { return static_cast<BoutReal>(jz) / static_cast<BoutReal>(GlobalNz); }

void Mesh::communicateFields(const std::vector<Field3D*>& fields) {
  const auto g = grid();
  std::vector<double*> ptrs;
  // A guard-cell exchange copies cells, so a field that was the difference of two others communicated in the same
  // call still is afterwards: carry such recipes over to the blocks' new versions.
  struct Kept {
    Field3D* f;
    Field3D* a;
    Field3D* b;
  };
  std::vector<Kept> kept;
  for (Field3D* f : fields) {
    requireData(*f, "Mesh::communicate");
    const auto& rec = f->block()->recipe;
    if (rec.op == 1) {
      Field3D *fa = nullptr, *fb = nullptr;
      for (Field3D* o : fields) {
        if (o->block() == rec.a && o->block()->version == rec.a_ver) {
          fa = o;
        }
        if (o->block() == rec.b && o->block()->version == rec.b_ver) {
          fb = o;
        }
      }
      if (fa != nullptr && fb != nullptr) {
        kept.push_back({f, fa, fb});
      }
    }
  }
  // ... and while such a difference is still a pending combination (never launched) it needs no exchange of its own:
  // evaluated later over the whole array from the exchanged a and b it has exactly the guard cells an exchange would
  // have given it.  It is taken off its operands' dependents for the duration, so that un-sharing them for the
  // exchange does not launch it.
  // Its handles on a and b are dropped for the duration as well (they would make the exchange un-share, i.e. copy,
  // both operands) and the combination is rebuilt from the same fields afterwards.
  std::vector<Field3D*> deferred;
  for (const Kept& k : kept) {
    auto& pend = k.f->block()->pending;
    const bool plain_difference = pend.size() == 2 && pend[0].coef == 1.0 && pend[1].coef == -1.0
                                  && pend[0].f1.get() == k.a->block() && pend[1].f1.get() == k.b->block() && !pend[0].f2
                                  && !pend[0].dv && !pend[1].f2 && !pend[1].dv && k.f != k.a && k.f != k.b;
    if (plain_difference) {
      deferred.push_back(k.f);
      pend.clear();
      for (Field3D* o : {k.a, k.b}) {
        auto& deps = o->block()->dependents;
        deps.erase(std::remove_if(deps.begin(), deps.end(),
                                  [&](const std::weak_ptr<bout::b200::Block>& w) { return w.lock().get() == k.f->block(); }),
                   deps.end());
      }
    }
  }
  auto isDeferred = [&](const Field3D* f) { return std::find(deferred.begin(), deferred.end(), f) != deferred.end(); };
  for (Field3D* f : fields) {
    if (!isDeferred(f)) {
      ptrs.push_back(f->wptr());
    }
  }
  for (const Kept& k : kept) {
    if (isDeferred(k.f)) {
      bout::b200::Block::LazyTerm ta, tb;
      ta.coef = 1.0;
      ta.f1 = k.a->shared();
      tb.coef = -1.0;
      tb.f1 = k.b->shared();
      k.f->block()->pending = {ta, tb};
      k.a->block()->dependents.push_back(k.f->shared());
      k.b->block()->dependents.push_back(k.f->shared());
    }
  }
  if (NYPE > 1) { // Y first, between ranks; then X (below) with the y guard rows just received (mesh.cxx:351-362)
    exchangeYDevice(g, MYG, ptrs);
  }
  if (NXPE > 1) {
    if (periodicY && MYG > 0 && NYPE == 1) { // Y wrap on this rank first
      for (size_t i = 0; i < ptrs.size(); i += 8) {
        const int n = static_cast<int>(std::min<size_t>(8, ptrs.size() - i));
        check(bout_b200_halo_self(&g, n, ptrs.data() + i, 0, 1, nullptr));
      }
    }
    exchangeXDevice(this, ptrs);
  }
  const int self_y = (periodicY && MYG > 0 && NYPE == 1 && NXPE == 1) ? 1 : 0;
  for (size_t i = 0; (self_y || periodicX) && i < ptrs.size(); i += 8) { // the kernel takes up to 8 fields per launch
    const int n = static_cast<int>(std::min<size_t>(8, ptrs.size() - i));
    check(bout_b200_halo_self(&g, n, ptrs.data() + i, periodicX ? 1 : 0, self_y, nullptr));
  }
  for (const Kept& k : kept) {
    bout::b200::Block::Recipe rec;
    rec.op = 1;
    rec.a = k.a->block();
    rec.b = k.b->block();
    rec.a_ver = rec.a->version;
    rec.b_ver = rec.b->version;
    k.f->block()->recipe = rec;
  }
}

namespace {
Field2D field2DFromOption(Mesh* m, Options& sec, const std::string& key, double def) {
  std::vector<double> v(static_cast<size_t>(m->LocalNx) * m->LocalNy, def);
  if (sec.isSet(key)) {
    Options& o = sec[key];
    for (int x = 0; x < m->LocalNx; ++x) {
      for (int y = 0; y < m->LocalNy; ++y) {
        v[static_cast<size_t>(x) * m->LocalNy + y] =
            bout::b200::evalOption(o, m->GlobalX(x), TWOPI * m->GlobalY(y), 0.0);
      }
    }
  }
  Field2D f(m);
  f.setHost(v);
  return f;
}
} // namespace

namespace {
// interpolateAndExtrapolate(f, CELL_CENTRE, extrapolate, extrapolate) on one rank (src/mesh/coordinates.cxx:32-125):
// communicate first (periodic directions wrap), then -- only where the caller extrapolates (the derived G1..G3; mesh
// quantities that come from options are valid in the boundary cells and are NOT extrapolated, :394-397) -- every
// boundary guard cell, outwards, from the three cells inside it (a copy with fewer than 3 interior cells).  X regions
// cover ystart..yend, Y regions xstart..xend (boutmesh.cxx:1028-1059 defaults).
void fixGuards2D(Mesh* m, std::vector<double>& f, bool extrapolate) {
  const int nx = m->LocalNx, ny = m->LocalNy;
  auto at = [&](int x, int y) -> double& { return f[static_cast<size_t>(x) * ny + y]; };
  const int mx = m->xend - m->xstart + 1, my = m->yend - m->ystart + 1;
  if (m->NYPE > 1) {
    m->exchangeY2D(f);
  } else if (m->periodicY && m->MYG > 0) { // Y first, then X with the y guards included (mesh.cxx:351-362)
    for (int x = 0; x < nx; ++x) {
      for (int i = 0; i < m->ystart; ++i) {
        at(x, i) = at(x, i + my);
        at(x, m->yend + 1 + i) = at(x, m->ystart + i);
      }
    }
  }
  if (m->NXPE > 1) {
    m->exchangeX2D(f);
  }
  if (m->periodicX && m->MXG > 0) {
    for (int y = 0; y < ny; ++y) {
      for (int i = 0; i < m->xstart; ++i) {
        at(i, y) = at(i + mx, y);
        at(m->xend + 1 + i, y) = at(m->xstart + i, y);
      }
    }
  }
  if (!extrapolate) {
    return; // options-defined quantities are valid in the boundary cells (GridFromOptions::hasXBoundaryGuards)
  }
  auto outward = [&](bool xdir, int first, int b, int width, int interior, int lo, int hi) {
    for (int o = lo; o <= hi; ++o) {
      for (int i = 0; i < width; ++i) {
        const int n = first + i * b;
        auto v = [&](int k) -> double& { return xdir ? at(k, o) : at(o, k); };
        v(n) = interior >= 3 ? 3.0 * v(n - b) - 3.0 * v(n - 2 * b) + v(n - 3 * b) : v(first - b);
      }
    }
  };
  if (!m->periodicX && m->MXG > 0 && m->firstX()) {
    outward(true, m->xstart - 1, -1, m->xstart, mx, m->ystart, m->yend);
  }
  if (!m->periodicX && m->MXG > 0 && m->lastX()) {
    outward(true, m->xend + 1, 1, nx - m->xend - 1, mx, m->ystart, m->yend);
  }
  if (!m->periodicY && m->MYG > 0) {
    outward(false, m->ystart - 1, -1, m->ystart, my, m->xstart, m->xend);
    outward(false, m->yend + 1, 1, ny - m->yend - 1, my, m->xstart, m->xend);
  }
}

// Index-space first derivative of a 2-D array along x or y with the mesh's default method, interior points only
// (index_derivs.cxx:70-78 C2 / C4), divided by the spacing (Coordinates::DDX / DDY, coordinates.cxx:1467-1490)
std::vector<double> ddIndex2D(Mesh* m, const std::vector<double>& f, const std::vector<double>& spacing, char dir) {
  const int nx = m->LocalNx, ny = m->LocalNy;
  std::vector<double> r(f.size(), 0.0);
  const int guards = dir == 'x' ? m->MXG : m->MYG;
  if (guards == 0) {
    return r; // no variation represented along this direction
  }
  const std::string method = m->method(dir, "first");
  const size_t st = dir == 'x' ? static_cast<size_t>(ny) : 1;
  if (method != "C2" && method != "C4") { // any consistent stencil gives 0 for a quantity that does not vary
    for (int x = m->xstart; x <= m->xend; ++x) {
      for (int y = m->ystart; y <= m->yend; ++y) {
        const size_t i = static_cast<size_t>(x) * ny + y;
        if (f[i + st] != f[i] || f[i - st] != f[i]) {
          throw BoutException("Coordinates: metric derivatives support mesh:dd{}:first = C2 or C4, not '{}'", dir, method);
        }
      }
    }
    return r;
  }
  const bool c4 = method == "C4" && guards >= 2;
  for (int x = m->xstart; x <= m->xend; ++x) {
    for (int y = m->ystart; y <= m->yend; ++y) {
      const size_t i = static_cast<size_t>(x) * ny + y;
      const double d = c4 ? (8. * f[i + st] - 8. * f[i - st] + f[i - 2 * st] - f[i + 2 * st]) / 12.
                          : 0.5 * (f[i + st] - f[i - st]);
      r[i] = d / spacing[i];
    }
  }
  (void)nx;
  return r;
}
} // namespace

Coordinates::Coordinates(Mesh* mesh) : localmesh(mesh) {
  Options& o = *mesh->options;
  dx = field2DFromOption(mesh, o, "dx", 1.0);
  dy = field2DFromOption(mesh, o, "dy", 1.0);
  // dz: mesh:dz, else from zperiod / ZMIN, ZMAX (coordinates.cxx:  zlength = (ZMAX - ZMIN) * 2 pi)
  if (o.isSet("dz")) {
    dz = field2DFromOption(mesh, o, "dz", 1.0);
  } else {
    Options& root = Options::root();
    double zmin = root.isSet("ZMIN") ? root["ZMIN"].as<double>() : 0.0;
    double zmax = root.isSet("ZMAX") ? root["ZMAX"].as<double>() : 1.0;
    if (root.isSet("zperiod")) {
      zmin = 0.0;
      zmax = 1.0 / root["zperiod"].as<double>();
    }
    dz = Field2D((zmax - zmin) * TWOPI / mesh->LocalNz, mesh);
  }
  g11 = field2DFromOption(mesh, o, "g11", 1.0);
  g22 = field2DFromOption(mesh, o, "g22", 1.0);
  g33 = field2DFromOption(mesh, o, "g33", 1.0);
  g12 = field2DFromOption(mesh, o, "g12", 0.0);
  g13 = field2DFromOption(mesh, o, "g13", 0.0);
  g23 = field2DFromOption(mesh, o, "g23", 0.0);
  G1 = field2DFromOption(mesh, o, "G1", 0.0);
  G2 = field2DFromOption(mesh, o, "G2", 0.0);
  G3 = field2DFromOption(mesh, o, "G3", 0.0);
  Bxy = field2DFromOption(mesh, o, "Bxy", 1.0);
  geometry();
}

int Coordinates::geometry() {
  // Covariant metric = inverse of the contravariant one, J = 1/sqrt(det g^ij) (coordinates.cxx:1210-1330),
  // d1_dx = -DDX(dx)/dx^2 in index space (coordinates.cxx:1088-1099), the Christoffel contractions G1..G3 and Bxy.
  // Set-up arithmetic on the host, once per mesh, as in the reference.
  Mesh* m = localmesh;
  auto a11 = g11.getHost(), a22 = g22.getHost(), a33 = g33.getHost(), a12 = g12.getHost(), a13 = g13.getHost(),
       a23 = g23.getHost(), hdx = dx.getHost(), hdy = dy.getHost(), hdz = dz.getHost();
  // mesh quantities read from options carry no boundary guard cells: the reference wraps / extrapolates them
  // (extrapolate_x / extrapolate_y default to true for such sources, coordinates.cxx:394-470)
  for (auto* a : {&a11, &a22, &a33, &a12, &a13, &a23, &hdx, &hdy}) {
    fixGuards2D(m, *a, false);
  }
  g11.setHost(a11), g22.setHost(a22), g33.setHost(a33), g12.setHost(a12), g13.setHost(a13), g23.setHost(a23);
  dx.setHost(hdx), dy.setHost(hdy);
  const size_t n = a11.size();
  std::vector<double> b11(n), b22(n), b33(n), b12(n), b13(n), b23(n), hJ(n), d1x(n, 0.0), d1y(n, 0.0);
  for (size_t i = 0; i < n; ++i) {
    const double det = a11[i] * (a22[i] * a33[i] - a23[i] * a23[i]) - a12[i] * (a12[i] * a33[i] - a23[i] * a13[i])
                       + a13[i] * (a12[i] * a23[i] - a22[i] * a13[i]);
    if (!(det > 0.0)) {
      throw BoutException("Coordinates: metric is not positive definite at index {}", i);
    }
    b11[i] = (a22[i] * a33[i] - a23[i] * a23[i]) / det;
    b22[i] = (a11[i] * a33[i] - a13[i] * a13[i]) / det;
    b33[i] = (a11[i] * a22[i] - a12[i] * a12[i]) / det;
    b12[i] = (a13[i] * a23[i] - a12[i] * a33[i]) / det;
    b13[i] = (a12[i] * a23[i] - a13[i] * a22[i]) / det;
    b23[i] = (a12[i] * a13[i] - a11[i] * a23[i]) / det;
    hJ[i] = 1.0 / std::sqrt(det);
  }
  // non-uniform corrections (coordinates.cxx:1085-1110, no d2x / d2y in the source): d1_dx = index-space DDX(1/dx),
  // d1_dy = DDY(1/dy) (sic: the coordinate derivative, i.e. divided by dy once more), both extrapolated into guards
  {
    std::vector<double> inv_dx(n), inv_dy(n), ones(n, 1.0);
    for (size_t i = 0; i < n; ++i) {
      inv_dx[i] = 1.0 / hdx[i];
      inv_dy[i] = 1.0 / hdy[i];
    }
    d1x = ddIndex2D(m, inv_dx, ones, 'x');
    d1y = ddIndex2D(m, inv_dy, hdy, 'y');
    fixGuards2D(m, d1x, true);
    fixGuards2D(m, d1y, true);
  }
  for (auto* a : {&b11, &b22, &b33, &b12, &b13, &b23, &hJ}) { // coordinates.cxx:500-530
    fixGuards2D(m, *a, false);
  }
  // Christoffel contractions G^i = (d_x(J g^i1) + d_y(J g^i2)) / J for z-independent metrics (coordinates.cxx:1018-1026,
  // then :1066-1068), unless the options give them; Bxy = sqrt(g_22) / J unless given (:547)
  auto contraction = [&](const std::vector<double>& gi1, const std::vector<double>& gi2) {
    std::vector<double> p(n), q(n);
    for (size_t i = 0; i < n; ++i) {
      p[i] = hJ[i] * gi1[i];
      q[i] = hJ[i] * gi2[i];
    }
    const auto dpx = ddIndex2D(m, p, hdx, 'x'), dqy = ddIndex2D(m, q, hdy, 'y');
    std::vector<double> G(n, 0.0);
    for (size_t i = 0; i < n; ++i) {
      G[i] = (dpx[i] + dqy[i]) / hJ[i];
    }
    fixGuards2D(m, G, true); // coordinates.cxx:1066-1068 extrapolate in x and y
    return G;
  };
  if (!localmesh->options->isSet("G1")) {
    G1.setHost(contraction(a11, a12));
  }
  if (!localmesh->options->isSet("G2")) {
    G2.setHost(contraction(a12, a22));
  }
  if (!localmesh->options->isSet("G3")) {
    G3.setHost(contraction(a13, a23));
  }
  if (!localmesh->options->isSet("Bxy")) {
    std::vector<double> hb(n);
    for (size_t i = 0; i < n; ++i) {
      hb[i] = std::sqrt(b22[i]) / hJ[i];
    }
    fixGuards2D(m, hb, false);
    Bxy.setHost(hb);
  }
  g_11.setHost(b11);
  g_22.setHost(b22);
  g_33.setHost(b33);
  g_12.setHost(b12);
  g_13.setHost(b13);
  g_23.setHost(b23);
  if (localmesh->options->isSet("J")) {
    J = field2DFromOption(m, *localmesh->options, "J", 1.0);
  } else {
    J.setHost(hJ);
  }
  d1_dx.setHost(d1x);
  d1_dy.setHost(d1y);
  non_uniform = localmesh->options->isSet("non_uniform") ? (*localmesh->options)["non_uniform"].as<bool>() : true;
  zlength_ = hdz[0] * m->LocalNz;
  cache.clear();
  return 0;
}

const Field2D& Coordinates::recip(const std::string& which) {
  // which metric fields the entry depends on
  const Field2D* src[2] = {nullptr, nullptr};
  if (which == "1/dx" || which == "1/dx2") { src[0] = &dx; }
  else if (which == "1/dy" || which == "1/dy2") { src[0] = &dy; }
  else if (which == "1/dz" || which == "1/dz2") { src[0] = &dz; }
  else if (which == "d1_dx/dx") { src[0] = &d1_dx; src[1] = &dx; }
  else if (which == "d1_dy/dy") { src[0] = &d1_dy; src[1] = &dy; }
  CacheKey key;
  for (int i = 0; i < 2; ++i) {
    if (src[i] != nullptr && src[i]->block() != nullptr) {
      key.blk[i] = src[i]->block();
      key.ver[i] = src[i]->block()->version;
    }
  }
  auto it = cache.find(which);
  if (it != cache.end()) {
    const CacheKey& k = cache_key[which];
    if (k.blk[0] == key.blk[0] && k.blk[1] == key.blk[1] && k.ver[0] == key.ver[0] && k.ver[1] == key.ver[1]) {
      return it->second;
    }
    cache.erase(it); // a source field was assigned or written since: rebuild
  }
  cache_key[which] = key;
  Field2D r(localmesh);
  if (which == "1/dx") { r = 1.0 / dx; }
  else if (which == "1/dy") { r = 1.0 / dy; }
  else if (which == "1/dz") { r = 1.0 / dz; }
  else if (which == "1/dx2") { r = 1.0 / (dx * dx); }
  else if (which == "1/dy2") { r = 1.0 / (dy * dy); }
  else if (which == "1/dz2") { r = 1.0 / (dz * dz); }
  else if (which == "d1_dx/dx") { r = d1_dx / dx; }
  else if (which == "d1_dy/dy") { r = d1_dy / dy; }
  else { throw BoutException("Coordinates::recip: unknown '{}'", which); }
  return cache.emplace(which, r).first->second;
}

// ======================================================================================================
// Derivatives, brackets, Delp2
// ======================================================================================================
namespace {
int regionId(const std::string& r) {
  if (r == "RGN_NOBNDRY") { return BOUT_B200_RGN_NOBNDRY; }
  if (r == "RGN_ALL") { return BOUT_B200_RGN_ALL; }
  if (r == "RGN_NOX") { return BOUT_B200_RGN_NOX; }
  if (r == "RGN_NOY") { return BOUT_B200_RGN_NOY; }
  throw BoutException("Unknown region '{}'", r);
}
int stencilOp(const std::string& family, const std::string& method) {
  static const std::map<std::string, int> table = {
      {"DDX_C2", BOUT_B200_DDX_C2}, {"DDX_C4", BOUT_B200_DDX_C4}, {"DDX_W2", BOUT_B200_DDX_W2},
      {"DDX_S2", BOUT_B200_DDX_S2}, {"DDX_W3", BOUT_B200_DDX_W3}, {"D2DX2_C2", BOUT_B200_D2DX2_C2},
      {"D2DX2_C4", BOUT_B200_D2DX2_C4}, {"D4DX4_C2", BOUT_B200_D4DX4_C2}, {"VDDX_C2", BOUT_B200_VDDX_C2},
      {"VDDX_C4", BOUT_B200_VDDX_C4}, {"VDDX_U1", BOUT_B200_VDDX_U1}, {"VDDX_U2", BOUT_B200_VDDX_U2},
      {"VDDX_U3", BOUT_B200_VDDX_U3}, {"VDDX_W3", BOUT_B200_VDDX_W3}, {"FDDX_U1", BOUT_B200_FDDX_U1},
      {"FDDX_U2", BOUT_B200_FDDX_U2}, {"FDDX_C2", BOUT_B200_FDDX_C2}, {"FDDX_C4", BOUT_B200_FDDX_C4}};
  const auto it = table.find(family + "_" + method);
  if (it == table.end()) {
    throw BoutException("Derivative method '{}' is not registered for {}", method, family);
  }
  return it->second;
}
std::string pick(Mesh* m, char dir, const std::string& kind, const std::string& method) {
  std::string s = method;
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::toupper(c); });
  return (s == "DEFAULT" || s.empty()) ? m->method(dir, kind) : s;
}
// one fused kernel: out = scale1 * op(f) [+ scale2 * op2(f)]
Field3D stencil(const Field3D& f, const Field3D* v, int op, int dir, const std::string& region, const Field2D* s1,
                int op2 = -1, const Field2D* s2 = nullptr) {
  requireData(f, "derivative");
  Field3D r(f.getMesh());
  const auto g = gridOf(f);
  check(bout_b200_stencil(&g, op, dir, regionId(region), v ? v->dptr() : nullptr, f.dptr(), nullptr, nullptr,
                          s1 ? s1->dptr() : nullptr, op2, s2 ? s2->dptr() : nullptr, r.wptr(), nullptr));
  return r;
}
} // namespace

Field3D DDX(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, nullptr, stencilOp("DDX", pick(m, 'x', "first", method)), BOUT_B200_X, region,
                 &m->getCoordinates()->recip("1/dx"));
}
Field3D DDY(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, nullptr, stencilOp("DDX", pick(m, 'y', "first", method)), BOUT_B200_Y, region,
                 &m->getCoordinates()->recip("1/dy"));
}
Field3D DDZ(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  const std::string meth = pick(m, 'z', "first", method);
  if (meth == "FFT") {
    Field3D r(m);
    const auto g = gridOf(f);
    check(bout_b200_ddz_fft(&g, 1, 0.0, regionId(region), f.dptr(), m->getCoordinates()->recip("1/dz").dptr(), r.wptr(), nullptr));
    return r;
  }
  return stencil(f, nullptr, stencilOp("DDX", meth), BOUT_B200_Z, region, &m->getCoordinates()->recip("1/dz"));
}
Field3D D2DX2(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  const int op = stencilOp("D2DX2", pick(m, 'x', "second", method));
  if (c->non_uniform) { // derivs.cxx:149-166: + d1_dx * DDX(f) / dx in the same pass
    return stencil(f, nullptr, op, BOUT_B200_X, region, &c->recip("1/dx2"),
                   stencilOp("DDX", m->method('x', "first")), &c->recip("d1_dx/dx"));
  }
  return stencil(f, nullptr, op, BOUT_B200_X, region, &c->recip("1/dx2"));
}
Field3D D2DY2(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  const int op = stencilOp("D2DX2", pick(m, 'y', "second", method));
  if (c->non_uniform) {
    return stencil(f, nullptr, op, BOUT_B200_Y, region, &c->recip("1/dy2"),
                   stencilOp("DDX", m->method('y', "first")), &c->recip("d1_dy/dy"));
  }
  return stencil(f, nullptr, op, BOUT_B200_Y, region, &c->recip("1/dy2"));
}
Field3D D2DZ2(const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  const std::string meth = pick(m, 'z', "second", method);
  if (meth == "FFT") {
    Field3D r(m);
    const auto g = gridOf(f);
    check(bout_b200_ddz_fft(&g, 2, 0.0, regionId(region), f.dptr(), m->getCoordinates()->recip("1/dz2").dptr(), r.wptr(), nullptr));
    return r;
  }
  return stencil(f, nullptr, stencilOp("D2DX2", meth), BOUT_B200_Z, region, &m->getCoordinates()->recip("1/dz2"));
}
Field3D VDDY(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("VDDX", pick(m, 'y', "upwind", method)), BOUT_B200_Y, region,
                 &m->getCoordinates()->recip("1/dy"));
}
Field3D FDDX(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("FDDX", pick(m, 'x', "flux", method)), BOUT_B200_X, region,
                 &m->getCoordinates()->recip("1/dx"));
}
Field3D FDDY(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("FDDX", pick(m, 'y', "flux", method)), BOUT_B200_Y, region,
                 &m->getCoordinates()->recip("1/dy"));
}
Field3D FDDZ(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("FDDX", pick(m, 'z', "flux", method)), BOUT_B200_Z, region,
                 &m->getCoordinates()->recip("1/dz"));
}

// ---- parallel operators, identity parallel transform (coordinates.cxx:1517-1625) -------------------------------------
const Field2D& Coordinates::invSg() {
  CacheKey key;
  key.blk[0] = g_22.block();
  key.ver[0] = g_22.block() != nullptr ? g_22.block()->version : 0;
  auto it = cache.find("invSg");
  if (it != cache.end()) {
    const CacheKey& k = cache_key["invSg"];
    if (k.blk[0] == key.blk[0] && k.ver[0] == key.ver[0]) {
      return it->second;
    }
    cache.erase(it);
  }
  cache_key["invSg"] = key;
  const std::vector<double> g = g_22.getHost();
  std::vector<double> v(g.size());
  for (size_t i = 0; i < g.size(); ++i) {
    v[i] = 1.0 / std::sqrt(g[i]);
  }
  Field2D r(localmesh);
  r.setHost(v);
  return cache.emplace("invSg", r).first->second;
}
const Field2D& Coordinates::Grad2_par2_DDY_invSg(const std::string& method) {
  const std::string name = "G2P2:" + method;
  CacheKey key;
  key.blk[0] = g_22.block();
  key.ver[0] = g_22.block() != nullptr ? g_22.block()->version : 0;
  key.blk[1] = dy.block();
  key.ver[1] = dy.block() != nullptr ? dy.block()->version : 0;
  auto it = cache.find(name);
  if (it != cache.end()) {
    const CacheKey& k = cache_key[name];
    if (k.blk[0] == key.blk[0] && k.ver[0] == key.ver[0] && k.blk[1] == key.blk[1] && k.ver[1] == key.ver[1]) {
      return it->second;
    }
    cache.erase(it);
  }
  cache_key[name] = key;
  // DDY of a Field2D over RGN_NOBNDRY with the mesh's first-derivative method, times invSg; the y guard cells of
  // invSg are those of the metric itself (wrapped where y is periodic: what the reference's communicate gives)
  const int nx = localmesh->LocalNx, ny = localmesh->LocalNy;
  const std::vector<double> isg = invSg().getHost(), d = dy.getHost();
  std::vector<double> v(isg.size(), 0.0);
  const bool c4 = method == "C4";
  if (!c4 && method != "C2") {
    throw BoutException("Grad2_par2: first-derivative method '{}' is not available for Field2D on the B200 host layer", method);
  }
  for (int x = 0; x < nx; ++x) {
    for (int y = localmesh->ystart; y <= localmesh->yend; ++y) {
      const size_t i = static_cast<size_t>(x) * ny + y;
      double dd;
      if (c4) {
        dd = (8.0 * isg[i + 1] - 8.0 * isg[i - 1] + isg[i - 2] - isg[i + 2]) / 12.0;
      } else {
        dd = 0.5 * (isg[i + 1] - isg[i - 1]);
      }
      v[i] = dd / d[i] * isg[i];
    }
  }
  Field2D r(localmesh);
  r.setHost(v);
  return cache.emplace(name, r).first->second;
}
Field3D Grad_par(const Field3D& f, CELL_LOC loc, const std::string& method) {
  return DDY(f, loc, method) * f.getMesh()->getCoordinates()->invSg();
}
Field3D Vpar_Grad_par(const Field3D& v, const Field3D& f, CELL_LOC loc, const std::string& method) {
  return VDDY(v, f, loc, method) * f.getMesh()->getCoordinates()->invSg();
}
Field3D Div_par(const Field3D& f, CELL_LOC loc, const std::string& method) {
  Coordinates* c = f.getMesh()->getCoordinates();
  return c->Bxy * Grad_par(f / c->Bxy, loc, method);
}
Field3D Grad2_par2(const Field3D& f, CELL_LOC loc, const std::string& method) {
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  const Field3D d1 = DDY(f, loc, method);
  const Field3D r2 = D2DY2(f, loc, method) / c->g_22;
  return c->Grad2_par2_DDY_invSg(pick(m, 'y', "first", method)) * d1 + r2;
}

Field3D Div_par_K_Grad_par(BoutReal kY, const Field3D& f, CELL_LOC loc) { return kY * Grad2_par2(f, loc); }
Field3D Div_par_K_Grad_par(const Field3D& kY, const Field3D& f, CELL_LOC loc) {
  return kY * Grad2_par2(f, loc) + Div_par(kY, loc) * Grad_par(f, loc);
}

Field3D VDDX(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("VDDX", pick(m, 'x', "upwind", method)), BOUT_B200_X, region,
                 &m->getCoordinates()->recip("1/dx"));
}
Field3D VDDZ(const Field3D& v, const Field3D& f, CELL_LOC, const std::string& method, const std::string& region) {
  Mesh* m = f.getMesh();
  return stencil(f, &v, stencilOp("VDDX", pick(m, 'z', "upwind", method)), BOUT_B200_Z, region,
                 &m->getCoordinates()->recip("1/dz"));
}

Field3D bracket(const Field3D& f, const Field3D& gfield, BRACKET_METHOD method, CELL_LOC, Solver*) {
  if (method == BRACKET_SIMPLE) { // difops.cxx:1042-1046: a subset of terms, built from the upwind operators
    return VDDX(DDZ(f), gfield) + VDDZ(-DDX(f), gfield);
  }
  if (method != BRACKET_ARAKAWA) {
    throw BoutException("bracket: BRACKET_ARAKAWA and BRACKET_SIMPLE are implemented in the B200 host layer (method {})",
                        static_cast<int>(method));
  }
  requireData(f, "bracket");
  requireData(gfield, "bracket");
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  Field3D r(m);
  const auto g = m->grid();
  check(bout_b200_bracket_arakawa(&g, f.dptr(), gfield.dptr(), c->dx.dptr(), c->dz.dptr(), r.wptr(), nullptr));
  return r;
}
Field3D bracket(const Field3D& f, const Field2D& gfield, BRACKET_METHOD method, CELL_LOC, Solver*) {
  if (method != BRACKET_ARAKAWA) {
    throw BoutException("bracket: only BRACKET_ARAKAWA is implemented on the B200 path");
  }
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  Field3D r(m);
  const auto g = m->grid();
  check(bout_b200_bracket_arakawa_3d2d(&g, f.dptr(), gfield.dptr(), c->dx.dptr(), c->dz.dptr(), r.wptr(), nullptr));
  return r;
}
Field3D bracket(const Field2D& f, const Field3D& g, BRACKET_METHOD method, CELL_LOC loc, Solver* s) {
  return -bracket(g, f, method, loc, s); // difops.cxx:796-799
}

Field3D Delp2(const Field3D& f, CELL_LOC, bool useFFT) {
  requireData(f, "Delp2");
  Mesh* m = f.getMesh();
  Coordinates* c = m->getCoordinates();
  Field3D r(m);
  const auto g = m->grid();
  if (useFFT) { // coordinates.cxx:1662-1699 with the default Laplacian instance's settings
    bout_b200_laplace L{};
    Laplacian::defaultInstance()->fill(L);
    check(bout_b200_delp2_fft(&g, &L, f.dptr(), r.wptr(), nullptr));
  } else {
    check(bout_b200_delp2_fd(&g, f.dptr(), c->dx.dptr(), c->dz.dptr(), c->g11.dptr(), c->g33.dptr(), c->g13.dptr(),
                             c->G1.dptr(), c->G3.dptr(), c->non_uniform ? c->d1_dx.dptr() : nullptr, r.wptr(), nullptr));
  }
  return r;
}

// ======================================================================================================
// LaplaceXZ (cyclic) and InvertPar (cyclic): SURVEY 8(f)4
// ======================================================================================================
LaplaceXZ::LaplaceXZ(Mesh* m, Options* options, CELL_LOC) : localmesh(m != nullptr ? m : bout::globals::mesh) {
  Options& o = options != nullptr ? *options : Options::root()["laplacexz"];
  const std::string type = o["type"].withDefault(std::string("cyclic"));
  if (type != "cyclic") {
    throw BoutException("LaplaceXZ type '{}' is not available: the B200 path implements 'cyclic'", type);
  }
  inner_boundary_flags = o["inner_boundary_flags"].withDefault(0);
  outer_boundary_flags = o["outer_boundary_flags"].withDefault(0);
  A = Field2D(1.0, localmesh); // laplacexz-cyclic.cxx:53-60
  B = Field2D(0.0, localmesh);
}
LaplaceXZ::~LaplaceXZ() {
  if (workspace != nullptr) {
    bout_b200_free(workspace);
  }
}
std::unique_ptr<LaplaceXZ> LaplaceXZ::create(Mesh* m, Options* opt, CELL_LOC loc) {
  return std::make_unique<LaplaceXZ>(m, opt, loc);
}
void LaplaceXZ::setCoefs(const Field2D& A2D, const Field2D& B2D) {
  A = A2D;
  B = B2D;
}
Field3D LaplaceXZ::solve(const Field3D& b, const Field3D& x0) {
  if (localmesh->NXPE > 1) {
    throw BoutException("LaplaceXZ: the x systems cross ranks; not available with NXPE > 1 in the B200 host layer");
  }
  requireData(b, "LaplaceXZ::solve");
  requireData(x0, "LaplaceXZ::solve");
  Coordinates* c = localmesh->getCoordinates();
  const auto g = localmesh->grid();
  bout_b200_laplacexz xz{};
  xz.A = A.dptr();
  xz.B = B.dptr();
  xz.J = c->J.dptr();
  xz.g11 = c->g11.dptr();
  xz.g33 = c->g33.dptr();
  xz.dx = c->dx.dptr();
  xz.zlength = c->zlength();
  xz.inner_boundary_flags = inner_boundary_flags;
  xz.outer_boundary_flags = outer_boundary_flags;
  const size_t need = bout_b200_laplacexz_workspace(&g);
  if (need > workspace_bytes) {
    if (workspace != nullptr) {
      bout_b200_free(workspace);
    }
    check(bout_b200_malloc(&workspace, need));
    workspace_bytes = need;
  }
  Field3D r(localmesh);
  check(bout_b200_laplacexz_solve(&g, &xz, b.dptr(), x0.dptr(), r.wptr(), workspace, workspace_bytes, nullptr));
  return r;
}

InvertPar::InvertPar(Options* options, CELL_LOC, Mesh* m) : localmesh(m != nullptr ? m : bout::globals::mesh) {
  Options& o = options != nullptr ? *options : Options::root()["parderiv"];
  const std::string type = o["type"].withDefault(std::string("cyclic"));
  if (type != "cyclic") {
    throw BoutException("InvertPar type '{}' is not available: the B200 path implements 'cyclic'", type);
  }
  if (localmesh->method('y', "first") != "C2") {
    throw BoutException("InvertPar: sg = DDY(1/sqrt(g_22))/sqrt(g_22) is formed with the C2 method on the B200 path "
                        "(mesh:ddy:first = {})", localmesh->method('y', "first"));
  }
  // parderiv/impls/cyclic/cyclic.cxx:55-56 on the host, once (set-up, as the reference's constructor)
  Coordinates* c = localmesh->getCoordinates();
  const std::vector<double> g22 = c->g_22.getHost(), dy = c->dy.getHost();
  const int nx = localmesh->LocalNx, ny = localmesh->LocalNy;
  std::vector<double> sgv(g22.size(), 0.0);
  for (int x = 0; x < nx; ++x) {
    for (int y = 1; y < ny - 1; ++y) {
      const size_t i = static_cast<size_t>(x) * ny + y;
      const double ddy = 0.5 * (1.0 / std::sqrt(g22[i + 1]) - 1.0 / std::sqrt(g22[i - 1])) / dy[i];
      sgv[i] = ddy / std::sqrt(g22[i]);
    }
  }
  sg = Field2D(localmesh);
  sg.setHost(sgv);
}
InvertPar::~InvertPar() {
  if (workspace != nullptr) {
    bout_b200_free(workspace);
  }
}
std::unique_ptr<InvertPar> InvertPar::create(Options* opt, CELL_LOC loc, Mesh* m) {
  return std::make_unique<InvertPar>(opt, loc, m);
}
Field3D InvertPar::solve(const Field3D& f) {
  requireData(f, "InvertPar::solve");
  Coordinates* c = localmesh->getCoordinates();
  const auto g = localmesh->grid();
  bout_b200_invert_par ip{};
  ip.A = hasA ? A.dptr() : nullptr;
  ip.B = hasB ? B.dptr() : nullptr;
  ip.C = hasC ? C.dptr() : nullptr;
  ip.D = hasD ? D.dptr() : nullptr;
  ip.E = hasE ? E.dptr() : nullptr;
  ip.sg = sg.dptr();
  ip.g_22 = c->g_22.dptr();
  ip.dy = c->dy.dptr();
  ip.d1_dy = c->non_uniform ? c->d1_dy.dptr() : nullptr;
  ip.twist_shift = nullptr; // the host mesh has no twist shift (identity parallel transform)
  ip.zlength = c->zlength();
  ip.closed = localmesh->periodicY ? 1 : 0;
  if (localmesh->NYPE > 1) {
    throw BoutException("InvertPar: field lines cross ranks; not available with NYPE > 1 in the B200 host layer");
  }
  const size_t need = bout_b200_invert_par_workspace(&g);
  if (need > workspace_bytes) {
    if (workspace != nullptr) {
      bout_b200_free(workspace);
    }
    check(bout_b200_malloc(&workspace, need));
    workspace_bytes = need;
  }
  Field3D r(localmesh);
  check(bout_b200_invert_par_solve(&g, &ip, f.dptr(), r.wptr(), workspace, workspace_bytes, nullptr));
  return r;
}

// ======================================================================================================
// Single-index loops (BOUT_FOR_RAJA): dispatch of a recorded loop body (include/bout/field_accessor.hxx)
// ======================================================================================================
namespace bout {
namespace b200 {
namespace si {
Recorder& Recorder::get() {
  static Recorder r;
  return r;
}

namespace {
struct Term { // coef * prim(f, g) [* prim2(f2, g2)]; prim = CONST means the constant coef itself
  double coef;
  Kind prim;
  const Field3D* f;
  const Field3D* g;
  Kind prim2 = CONST; // CONST = no second factor; else the term is the PRODUCT of two primitives (n[i] * phi[i] ...)
  const Field3D* f2 = nullptr;
  const Field3D* g2 = nullptr;
};

bool constantValue(const std::vector<Node>& nodes, int id, double* v) {
  const Node& n = nodes[id];
  double a = 0.0, b = 0.0;
  switch (n.kind) {
  case CONST: *v = n.value; return true;
  case NEG:
    if (!constantValue(nodes, n.a, &a)) {
      return false;
    }
    *v = -a;
    return true;
  case ADD:
  case SUB:
  case MUL:
  case DIV:
    if (!constantValue(nodes, n.a, &a) || !constantValue(nodes, n.b, &b)) {
      return false;
    }
    *v = n.kind == ADD ? a + b : (n.kind == SUB ? a - b : (n.kind == MUL ? a * b : a / b));
    return true;
  default: return false;
  }
}

// the expression as a sum of coef * primitive terms; products of two field-dependent factors are not linear
void linearise(const std::vector<Node>& nodes, int id, double scale, std::vector<Term>& out) {
  const Node& n = nodes[id];
  double c = 0.0;
  switch (n.kind) {
  case CONST: {
    Term t{};
    t.coef = scale * n.value;
    t.prim = CONST;
    t.f = t.g = nullptr;
    out.push_back(t);
    return;
  }
  case ADD:
    linearise(nodes, n.a, scale, out);
    linearise(nodes, n.b, scale, out);
    return;
  case SUB:
    linearise(nodes, n.a, scale, out);
    linearise(nodes, n.b, -scale, out);
    return;
  case NEG: linearise(nodes, n.a, -scale, out); return;
  case MUL:
    if (constantValue(nodes, n.a, &c)) {
      linearise(nodes, n.b, scale * c, out);
    } else if (constantValue(nodes, n.b, &c)) {
      linearise(nodes, n.a, scale * c, out);
    } else {
      // two field-dependent factors: each must be ONE primitive (times a constant), their product becomes one term
      std::vector<Term> ta, tb;
      linearise(nodes, n.a, 1.0, ta);
      linearise(nodes, n.b, 1.0, tb);
      if (ta.size() != 1 || tb.size() != 1 || ta[0].prim == CONST || tb[0].prim == CONST || ta[0].prim2 != CONST
          || tb[0].prim2 != CONST) {
        throw BoutException("BOUT_FOR_RAJA: only products of two single operator terms are implemented in a single-index "
                            "loop on the B200 path");
      }
      Term t = ta[0];
      t.coef = scale * ta[0].coef * tb[0].coef;
      t.prim2 = tb[0].prim;
      t.f2 = tb[0].f;
      t.g2 = tb[0].g;
      out.push_back(t);
    }
    return;
  case DIV:
    if (!constantValue(nodes, n.b, &c)) {
      throw BoutException("BOUT_FOR_RAJA: division by a field term in a single-index loop is not implemented on the B200 path");
    }
    linearise(nodes, n.a, scale / c, out);
    return;
  default: {
    Term t{};
    t.coef = scale;
    t.prim = n.kind;
    t.f = n.f;
    t.g = n.g;
    out.push_back(t);
    return;
  }
  }
}

// same primitive on the same fields: coefficients add
std::vector<Term> merged(const std::vector<Term>& in) {
  std::vector<Term> out;
  for (const Term& t : in) {
    bool found = false;
    for (Term& o : out) {
      if (o.prim == t.prim && o.f == t.f && o.g == t.g && o.prim2 == t.prim2 && o.f2 == t.f2 && o.g2 == t.g2) {
        o.coef += t.coef;
        found = true;
        break;
      }
    }
    if (!found) {
      out.push_back(t);
    }
  }
  return out;
}

const Term* findTerm(const std::vector<Term>& ts, Kind prim, const Field3D* f = nullptr, const Field3D* g = nullptr) {
  for (const Term& t : ts) {
    if (t.prim2 == CONST && t.prim == prim && (f == nullptr || t.f == f) && (g == nullptr || t.g == g)) {
      return &t;
    }
  }
  return nullptr;
}

// is x's block the difference a - b, still valid?
bool isDifference(const Field3D* x, const Field3D* a, const Field3D* b) {
  const Block* xb = x->block();
  if (xb == nullptr || a->block() == nullptr || b->block() == nullptr) {
    return false;
  }
  const Block::Recipe& r = xb->recipe;
  return r.op == 1 && r.a == a->block() && r.a_ver == a->block()->version && r.b == b->block()
         && r.b_ver == b->block()->version;
}

bout_b200_metric metricOf(Coordinates* c) {
  bout_b200_metric m{};
  m.dx = c->dx.dptr();
  m.dy = c->dy.dptr();
  m.dz = c->dz.dptr();
  m.J = c->J.dptr();
  m.g_22 = c->g_22.dptr();
  m.g11 = c->g11.dptr();
  m.g33 = c->g33.dptr();
  m.g13 = c->g13.dptr();
  m.G1 = c->G1.dptr();
  m.G3 = c->G3.dptr();
  m.d1_dx = c->non_uniform ? c->d1_dx.dptr() : nullptr;
  return m;
}

// examples/hasegawa-wakatani-3d/hw.cxx:58-74:
//   ddt(n)    = -[phi, n]    - alpha Div_par_Grad_par(phi - n) - kappa DDZ(phi) + Dn Delp2(n)
//   ddt(vort) = -[phi, vort] - alpha Div_par_Grad_par(phi - n)                  + Dvort Delp2(vort)
// Recognised from the recorded terms, whatever the order or grouping they were written in.
bool tryHw3d(Mesh* mesh, const std::vector<std::pair<Field3D*, std::vector<Term>>>& eqs) {
  if (eqs.size() != 2) {
    return false;
  }
  for (int swap = 0; swap < 2; ++swap) {
    Field3D* n = eqs[swap].first;
    Field3D* vort = eqs[1 - swap].first;
    const std::vector<Term>& tn = eqs[swap].second;
    const std::vector<Term>& tv = eqs[1 - swap].second;
    const Term* bn = findTerm(tn, BRACKET, nullptr, n);
    const Term* bv = findTerm(tv, BRACKET, nullptr, vort);
    if (bn == nullptr || bv == nullptr || bn->f != bv->f || bn->coef != -1.0 || bv->coef != -1.0) {
      continue;
    }
    const Field3D* phi = bn->f;
    const Term* dn = findTerm(tn, DIV_PAR_GRAD_PAR);
    const Term* dv = findTerm(tv, DIV_PAR_GRAD_PAR);
    if (dn == nullptr || dv == nullptr || dn->f != dv->f || dn->coef != dv->coef || !isDifference(dn->f, phi, n)) {
      continue;
    }
    const Term* zn = findTerm(tn, DDZ, phi);
    const Term* ln = findTerm(tn, DELP2, n);
    const Term* lv = findTerm(tv, DELP2, vort);
    const size_t used_n = 2 + (zn != nullptr) + (ln != nullptr), used_v = 2 + (lv != nullptr);
    if (tn.size() != used_n || tv.size() != used_v) {
      continue; // some other term is present
    }
    // scratch for the per-(x,y) coefficient table of the fused kernel (rebuilt from the metric by every call)
    const auto g = mesh->grid();
    const size_t need = bout_b200_fused_coef_doubles(&g);
    static std::map<Mesh*, std::shared_ptr<Block>> ws_blocks;
    std::shared_ptr<Block>& wb = ws_blocks[mesh];
    if (!wb || wb->count < need) {
      wb = acquire(need);
    }
    const bout_b200_metric m = metricOf(mesh->getCoordinates());
    Field3D* dtn = n->timeDeriv();
    Field3D* dtv = vort->timeDeriv();
    if (!dtn->isAllocated()) {
      dtn->allocate();
    }
    if (!dtv->isAllocated()) {
      dtv->allocate();
    }
    // a stage update the solver asked for (StageFusion) becomes the kernel's epilogue when it is about these two fields
    bout_b200_rk_epilogue ep{};
    bool fuse = false;
    StageFusion& sf = stageFusion();
    if (sf.active && !sf.consumed && sf.vars.size() == 2) {
      int idx[2] = {-1, -1};
      for (int e = 0; e < 2; ++e) {
        if (sf.vars[e] == n) {
          idx[0] = e;
        }
        if (sf.vars[e] == vort) {
          idx[1] = e;
        }
      }
      if (idx[0] >= 0 && idx[1] >= 0) {
        fuse = true;
        for (int e = 0; e < 2; ++e) {
          const int j = idx[e];
          ep.start[e] = sf.start[j] != nullptr ? sf.start[j]->dptr() : nullptr;
          ep.acc_in[e] = sf.acc_in[j] != nullptr ? sf.acc_in[j]->dptr() : nullptr;
        }
        for (int e = 0; e < 2; ++e) { // (write pointers last: acc_out may be the block acc_in / start read from)
          const int j = idx[e];
          ep.state_out[e] = sf.state_out[j] != nullptr ? sf.state_out[j]->wptr() : nullptr;
          ep.acc_out[e] = sf.acc_out[j] != nullptr ? sf.acc_out[j]->wptr() : nullptr;
        }
        ep.a = sf.a;
        ep.b = sf.b;
      }
    }
    check(bout_b200_hw3d_rhs(&g, &m, -dn->coef, zn != nullptr ? -zn->coef : 0.0, ln != nullptr ? ln->coef : 0.0,
                             lv != nullptr ? lv->coef : 0.0, n->dptr(), vort->dptr(), phi->dptr(),
                             fuse ? nullptr : dtn->wptr(), fuse ? nullptr : dtv->wptr(), fuse ? &ep : nullptr, wb->ptr,
                             nullptr));
    if (fuse) {
      sf.consumed = true;
      sf.ddt_blocks = {dtn->block(), dtv->block()};
      sf.ddt_versions = {dtn->block()->version, dtv->block()->version};
      sf.vars = {n, vort};
    }
    return true;
  }
  return false;
}

// one kernel per operator: every primitive becomes a field, the assignment a linear combination over the region
Field3D onePrimitive(Kind prim, const Field3D* f, const Field3D* g);
Field3D primitiveField(const Term& t) {
  if (t.prim2 != CONST) {
    return onePrimitive(t.prim, t.f, t.g) * onePrimitive(t.prim2, t.f2, t.g2);
  }
  return onePrimitive(t.prim, t.f, t.g);
}
Field3D onePrimitive(Kind prim, const Field3D* f, const Field3D* g) {
  struct {
    Kind prim;
    const Field3D* f;
    const Field3D* g;
  } t{prim, f, g};
  switch (t.prim) {
  case FIELD: return *t.f;
  case BRACKET: return ::bracket(*t.f, *t.g, BRACKET_ARAKAWA);
  case DDX: return ::DDX(*t.f, CELL_DEFAULT, "C2");
  case DDZ: return ::DDZ(*t.f, CELL_DEFAULT, "C2");
  case DELP2: return ::Delp2(*t.f, CELL_DEFAULT, false);
  default:
    throw BoutException("BOUT_FOR_RAJA: this single-index operator is only implemented inside the fused "
                        "hasegawa-wakatani-3d pattern on the B200 path (operator {})", static_cast<int>(t.prim));
  }
}
} // namespace

void dispatch(const Region<Ind3D>& region) {
  Recorder& rec = Recorder::get();
  if (rec.assigns.empty()) {
    return;
  }
  Mesh* mesh = rec.assigns.front().target->getMesh();
  if (&region != &mesh->getRegion3D("RGN_NOBNDRY")) {
    throw BoutException("BOUT_FOR_RAJA: only RGN_NOBNDRY loops are implemented on the B200 path");
  }
  // one term list per assigned field, in order of first assignment (a later plain assignment replaces, += adds)
  std::vector<std::pair<Field3D*, std::vector<Term>>> eqs;
  for (const Assignment& a : rec.assigns) {
    std::vector<Term> ts;
    linearise(rec.nodes, a.node, 1.0, ts);
    auto it = eqs.begin();
    for (; it != eqs.end(); ++it) {
      if (it->first == a.target) {
        break;
      }
    }
    if (it == eqs.end()) {
      if (a.accumulate) { // `ddt(f)[i] += ...` in a loop of its own: on top of what an earlier loop (or statement) left
        Field3D* cur = a.target->timeDeriv();
        if (!cur->isAllocated()) {
          throw BoutException("BOUT_FOR_RAJA: ddt(f)[i] += ... before any assignment to it");
        }
        Term t0{};
        t0.coef = 1.0;
        t0.prim = FIELD;
        t0.f = cur;
        t0.g = nullptr;
        ts.insert(ts.begin(), t0);
      }
      eqs.push_back({a.target, ts});
    } else if (a.accumulate) {
      it->second.insert(it->second.end(), ts.begin(), ts.end());
    } else {
      it->second = ts;
    }
  }
  for (auto& e : eqs) {
    e.second = merged(e.second);
  }
  if (tryHw3d(mesh, eqs)) {
    return;
  }
  const auto g = mesh->grid();
  for (auto& e : eqs) {
    std::vector<Field3D> held;
    std::vector<double> coef;
    double constant = 0.0;
    for (const Term& t : e.second) {
      if (t.prim == CONST) {
        constant += t.coef;
      } else {
        held.push_back(primitiveField(t));
        coef.push_back(t.coef);
      }
    }
    if (constant != 0.0) {
      held.push_back(Field3D(constant, mesh));
      coef.push_back(1.0);
    }
    Field3D* dt = e.first->timeDeriv();
    if (!dt->isAllocated()) {
      dt->allocate();
    }
    if (held.empty()) {
      *dt = 0.0;
      continue;
    }
    if (held.size() > 8) {
      throw BoutException("BOUT_FOR_RAJA: more than 8 terms in one assignment");
    }
    std::vector<const double*> x;
    for (const Field3D& f : held) {
      x.push_back(f.dptr());
    }
    check(bout_b200_lincomb(&g, BOUT_B200_RGN_NOBNDRY, static_cast<int>(coef.size()), coef.data(), x.data(), dt->wptr(),
                            nullptr));
  }
}
} // namespace si
} // namespace b200
} // namespace bout

// ======================================================================================================
// Laplacian
// ======================================================================================================
std::unique_ptr<Laplacian> Laplacian::instance;

Laplacian::Laplacian(Options* options, CELL_LOC, Mesh* mesh)
    : localmesh(mesh != nullptr ? mesh : bout::globals::mesh), coords(localmesh->getCoordinates()) {
  Options& o = options != nullptr ? *options : Options::root()["laplace"];
  std::string type = o["type"].withDefault(std::string("cyclic"));
  // pcr and pcr_thomas solve LaplaceCyclic's systems (same FFTTransform and tridagMatrix) with another elimination
  // order across MPI ranks (src/invert/laplace/impls/pcr/pcr.cxx, impls/pcr_thomas/pcr_thomas.cxx): one device path
  if (type != "cyclic" && type != "pcr" && type != "pcr_thomas") {
    throw BoutException("Laplacian type '{}' is not available: the B200 path implements 'cyclic', 'pcr' and 'pcr_thomas'", type);
  }
  pcr_thomas = type == "pcr_thomas";
  // invert_laplace.cxx:68-138
  const BoutReal filter = o["filter"].withDefault(0.0);
  const int ncz = localmesh->LocalNz;
  maxmode = o["maxmode"].withDefault(static_cast<int>(std::lround((1.0 - filter) * static_cast<BoutReal>(ncz) / 2.0)));
  maxmode = std::max(0, std::min(maxmode, ncz / 2));
  nonuniform = o["nonuniform"].withDefault(coords->non_uniform);
  all_terms = o["all_terms"].withDefault(true);
  global_flags = o["global_flags"].withDefault(0);
  inner_boundary_flags = o["inner_boundary_flags"].withDefault(0);
  outer_boundary_flags = o["outer_boundary_flags"].withDefault(0);
  include_yguards = o["include_yguards"].withDefault(false);
}

Laplacian::~Laplacian() {
  if (workspace != nullptr) {
    bout_b200_free(workspace);
  }
  if (workspace_perp != nullptr) {
    bout_b200_free(workspace_perp);
  }
}

std::unique_ptr<Laplacian> Laplacian::create(Options* opts, CELL_LOC loc, Mesh* mesh, Solver*) {
  return std::make_unique<Laplacian>(opts, loc, mesh);
}
Laplacian* Laplacian::defaultInstance() {
  if (!instance) {
    instance = create();
  }
  return instance.get();
}
void Laplacian::cleanup() { instance.reset(); }

void Laplacian::fill(bout_b200_laplace& L) const {
  L = bout_b200_laplace{};
  L.dx = coords->dx.dptr();
  L.g11 = coords->g11.dptr();
  L.g33 = coords->g33.dptr();
  L.g13 = coords->g13.dptr();
  L.G1 = coords->G1.dptr();
  L.G3 = coords->G3.dptr();
  L.g_11 = coords->g_11.dptr();
  L.g_22 = coords->g_22.dptr();
  L.A = hasA ? A.dptr() : nullptr;
  L.C1 = hasC ? C1.dptr() : nullptr;
  L.C2 = hasC ? C2.dptr() : nullptr;
  L.D = hasD ? D.dptr() : nullptr;
  L.zlength = coords->zlength();
  L.all_terms = all_terms ? 1 : 0;
  L.nonuniform = nonuniform ? 1 : 0;
  L.global_flags = global_flags;
  L.inner_boundary_flags = inner_boundary_flags;
  L.outer_boundary_flags = outer_boundary_flags;
  L.maxmode = maxmode;
  L.periodic_x = localmesh->periodicX ? 1 : 0;
  L.first_x = localmesh->firstX() ? 1 : 0;
  L.last_x = localmesh->lastX() ? 1 : 0;
  // cyclic_laplace.cxx:216-232: y guards are only solved where a y boundary exists
  L.include_yguards_lower = (include_yguards && !localmesh->periodicY) ? 1 : 0;
  L.include_yguards_upper = L.include_yguards_lower;
}

Field3D Laplacian::solve(const Field3D& b, const Field3D& x0) {
  if (pcr_thomas && localmesh->xstart > 1
      && ((global_flags & INVERT_BOTH_BNDRY_ONE) || ((inner_boundary_flags | outer_boundary_flags) & INVERT_BNDRY_ONE))) {
    // the reference's pcr_thomas leaves the outer guard rows out of its boundary elimination in this case and then
    // disagrees with its own cyclic and pcr solvers (tests/golden/make_golden_invert.py probes it): not reproduced
    throw BoutException("Laplacian type pcr_thomas with INVERT_BNDRY_ONE and more than one x guard cell is not implemented "
                        "on the B200 path; use cyclic or pcr");
  }
  requireData(b, "Laplacian::solve");
  const auto g = localmesh->grid();
  bout_b200_laplace L{};
  fill(L);
  const size_t need = bout_b200_laplace_cyclic_workspace(&g, &L);
  if (need > workspace_bytes) {
    if (workspace != nullptr) {
      check(bout_b200_free(workspace));
    }
    check(bout_b200_malloc(&workspace, need));
    workspace_bytes = need;
  }
  Field3D x(localmesh);
  // cells the solve does not write keep the initial guess (cyclic_laplace.cxx:196 "x = x0" semantics
  // for boundary/guard cells are emptyFrom in the reference; zero-initialised pool blocks here)
  if (localmesh->NXPE > 1) {
    // the tridiagonal systems span the ranks of the X row: local elimination, one all-gather of the interface rows,
    // interface system + back substitution (cyclic_reduction.hxx:169-440 without its gather / scatter of whole systems)
    const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    const size_t n = bout_b200_laplace_cyclic_iface_doubles(&g, &L);
    if (n > yd.capacity) {
      throw BoutException("Laplacian::solve with NXPE = {}: the interface record ({} doubles) exceeds the exchange buffers ({})",
                          yd.nxpe, n, yd.capacity);
    }
    check(bout_b200_laplace_cyclic_reduce(&g, &L, b.dptr(), x0.isAllocated() ? x0.dptr() : nullptr, workspace, workspace_bytes,
                                          yd.gather_send, nullptr));
    if (yd.allgather(n, yd.user) != 0) {
      throw BoutException("Laplacian::solve: the all-gather of the interface rows failed (rank {} of {})", yd.xind, yd.nxpe);
    }
    check(bout_b200_laplace_cyclic_finish(&g, &L, yd.gather_recv, yd.nxpe, yd.xind, x.wptr(), workspace, workspace_bytes,
                                          nullptr));
    return x;
  }
  check(bout_b200_laplace_cyclic_solve(&g, &L, b.dptr(), x0.isAllocated() ? x0.dptr() : nullptr, x.wptr(), workspace,
                                       workspace_bytes, nullptr));
  return x;
}

FieldPerp Laplacian::solve(const FieldPerp& b, const FieldPerp& x0) {
  if (localmesh->NXPE > 1) {
    throw BoutException("Laplacian::solve(FieldPerp) is not available with NXPE > 1 in the B200 host layer");
  }
  requireData(b, "Laplacian::solve(FieldPerp)");
  const int jy = b.getIndex();
  if (jy < 0 || jy >= localmesh->LocalNy) {
    throw BoutException("Laplacian::solve(FieldPerp): the FieldPerp has no y index");
  }
  const auto g = localmesh->grid();
  bout_b200_laplace L{};
  fill(L);
  const size_t need = bout_b200_laplace_cyclic_workspace_perp(&g, &L);
  if (need > workspace_perp_bytes) {
    if (workspace_perp != nullptr) {
      check(bout_b200_free(workspace_perp));
    }
    check(bout_b200_malloc(&workspace_perp, need));
    workspace_perp_bytes = need;
  }
  FieldPerp x(localmesh, CELL_CENTRE, jy);
  check(bout_b200_laplace_cyclic_solve_perp(&g, &L, jy, b.dptr(), x0.isAllocated() ? x0.dptr() : nullptr, x.wptr(),
                                            workspace_perp, workspace_perp_bytes, nullptr));
  return x;
}

// ======================================================================================================
// Solver: explicit integrators on device-resident fields (rk4, rk3ssp; euler, rkgeneric, splitrk, adams-bashforth below)
// ======================================================================================================
Solver::Solver(Options* opts) : options(opts != nullptr ? opts : &Options::root()["solver"]) {
  Options& root = Options::root();
  nout = (*options)["nout"].withDefault(root.isSet("nout") ? root["nout"].as<int>() : 1);
  output_timestep = (*options)["output_step"].withDefault(root.isSet("timestep") ? root["timestep"].as<double>() : 1.0);
  timestep = (*options)["timestep"].withDefault(output_timestep); // rk4.cxx:15-18
  is_nonsplit_model_diffusive = (*options)["is_nonsplit_model_diffusive"].withDefault(true); // solver.cxx:97-101
}

namespace {
void initialProfile(const std::string& name, Field3D& var) { // initialprofiles.cxx:33-48 (+ VAROPTION fallback to [all])
  Mesh* m = var.getMesh();
  Options& vo = Options::root()[name];
  Options& all = Options::root()["all"];
  Options* fo = vo.isSet("function") ? &vo["function"] : (all.isSet("function") ? &all["function"] : nullptr);
  double scale = 1.0;
  if (vo.isSet("scale")) {
    scale = vo["scale"].as<double>();
  } else if (all.isSet("scale")) {
    scale = all["scale"].as<double>();
  }
  std::vector<double> h(var.size(), 0.0);
  if (fo != nullptr && scale != 0.0) {
    bout::b200::forEachPlane(m->LocalNx, [&](int x) {
      for (int y = 0; y < m->LocalNy; ++y) {
        for (int z = 0; z < m->LocalNz; ++z) {
          h[(static_cast<size_t>(x) * m->LocalNy + y) * m->LocalNz + z] =
              scale * bout::b200::evalOption(*fo, m->GlobalX(x), TWOPI * m->GlobalY(y), TWOPI * m->GlobalZ(z));
        }
      }
    });
  }
  var.setHost(h);
}
} // namespace

void initial_profile(const std::string& name, Field3D& var) {
  if (var.getMesh() == nullptr) {
    var = Field3D(bout::globals::mesh);
  }
  initialProfile(name, var);
}

void Solver::add(Field3D& v, const std::string& name) {
  if (initialised) {
    throw BoutException("Error: Cannot add to solver after initialisation\n");
  }
  v.name = name;
  initialProfile(name, v);
  v.setBoundary(name);
  v.applyBoundary(true);
  Field3D* d = v.timeDeriv();
  *d = 0.0;
  f3d.push_back({&v, d, name});
}

void Solver::setModel(PhysicsModel* m) {
  model = m;
  m->initialise(this);
}

void PhysicsModel::initialise(Solver* s) {
  solver = s;
  mesh = bout::globals::mesh;
  if (init(false) != 0) {
    throw BoutException("Couldn't initialise physics model");
  }
}

void Solver::pre_rhs() {
  // solver.cxx:1486-1500 applies every evolving field's conditions side by side; here all x sides go in ONE launch and
  // all y sides (whose corner cells read what the x sides wrote) in a second one
  if (f3d.empty()) {
    return;
  }
  Mesh* m = f3d[0].var->getMesh();
  const auto g = m->grid();
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<bout_b200_bndry_spec> specs;
    for (auto& f : f3d) {
      const auto& conds = f.var->bndry_conditions;
      for (int s = 2 * pass; s < 2 * pass + 2 && s < static_cast<int>(conds.size()); ++s) {
        bout_b200_bndry_spec q;
        if (hasBoundary(m, s) && boundarySpec(*f.var, conds[s], s, q)) {
          specs.push_back(q);
        }
      }
    }
    if (!specs.empty()) {
      check(bout_b200_boundary_batch(&g, static_cast<int>(specs.size()), specs.data(), nullptr));
    }
  }
}
void Solver::post_rhs() {
  for (auto& f : f3d) {
    if (!f.F_var->isAllocated()) {
      throw BoutException("Time derivative for variable '{}' not set", f.name);
    }
  }
}

int Solver::run_rhs(BoutReal t) {
  int status;
  if (model->splitOperator()) { // solver.cxx:1391-1413 without the flat-array copies
    pre_rhs();
    status = model->runConvective(t);
    post_rhs();
    std::vector<Field3D> conv;
    for (auto& f : f3d) {
      conv.push_back(*f.F_var);
    }
    pre_rhs();
    status = model->runDiffusive(t);
    post_rhs();
    for (size_t i = 0; i < f3d.size(); ++i) {
      *f3d[i].F_var = conv[i] + *f3d[i].F_var;
    }
  } else {
    pre_rhs();
    status = model->runRHS(t);
    post_rhs();
  }
  ++rhs_calls;
  return status;
}

int Solver::run_convective(BoutReal t) { // solver.cxx:1430-1458
  int status = 0;
  pre_rhs();
  if (model->splitOperator()) {
    status = model->runConvective(t);
  } else if (!is_nonsplit_model_diffusive) {
    status = model->runRHS(t);
  } else {
    for (auto& f : f3d) {
      *f.F_var = 0.0;
    }
  }
  post_rhs();
  ++rhs_calls;
  return status;
}

int Solver::run_diffusive(BoutReal t) { // solver.cxx:1460-1484
  int status = 0;
  pre_rhs();
  if (model->splitOperator()) {
    status = model->runDiffusive(t);
    post_rhs();
  } else if (is_nonsplit_model_diffusive) {
    status = model->runRHS(t);
  } else {
    for (auto& f : f3d) {
      *f.F_var = 0.0;
    }
  }
  ++rhs_calls_i;
  return status;
}

void Solver::addMonitor(Monitor* monitor, MonitorPosition pos) {
  if (monitor->getTimestep() > 0 && monitor->getTimestep() != output_timestep) {
    throw BoutException("Solver::addMonitor: monitors with their own timestep are not supported by the B200 host layer");
  }
  if (pos == MonitorPosition::FRONT) {
    monitors.insert(monitors.begin(), monitor);
  } else {
    monitors.push_back(monitor);
  }
}
void Solver::removeMonitor(Monitor* monitor) {
  monitors.erase(std::remove(monitors.begin(), monitors.end(), monitor), monitors.end());
}
void Solver::call_monitors(BoutReal t, int iter, int nout_total) { // solver.cxx:874-935
  try {
    for (Monitor* m : monitors) {
      const int ret = m->call(this, t, iter, nout_total);
      if (ret != 0) {
        throw BoutException("Monitor signalled to quit (return code {})", ret);
      }
    }
    const int ret = model->outputMonitor(t, iter, nout_total); // PhysicsModelMonitor::call, physicsmodel.cxx:255
    if (ret != 0) {
      throw BoutException("Monitor signalled to quit (return code {})", ret);
    }
  } catch (const BoutException&) {
    for (Monitor* m : monitors) {
      m->cleanup();
    }
    throw;
  }
  if (iter == nout_total) {
    for (Monitor* m : monitors) {
      m->cleanup();
    }
  }
}

void Solver::take_step(BoutReal, BoutReal) {
  throw BoutException("Solver::take_step: this integrator advances through run_output_step only");
}

void Solver::run_output_step(BoutReal target) { // rk4.cxx:75-152 (non-adaptive), rk3-ssp.cxx:66-93
  bool running = true;
  do {
    BoutReal dt = timestep;
    running = true;
    if ((simtime + dt) >= target) {
      dt = target - simtime;
      running = false;
    }
    take_step(simtime, dt);
    simtime += dt;
  } while (running);
}

int Solver::solve(int nout_in, BoutReal output_step) {
  if (nout_in >= 0) {
    nout = nout_in;
    output_timestep = output_step;
  }
  if (initialised) {
    throw BoutException("ERROR: Solver is already initialised\n");
  }
  if (f3d.empty()) {
    throw BoutException("Solver::solve: no evolving variables were added");
  }
  initialised = true;
  // Restart (`restart = true`, file `[restart] path`): evolving variables, time and iteration of the latest output of
  // an earlier run (reference: Solver::add reads them from the restart options, solver.cxx:240-262; tt / hist_hi :694-697)
  Options& ro = Options::root()["restart_file"];
  // On a decomposed mesh every rank has files of its own, named like the reference's BOUT.dmp.<rank>.nc /
  // BOUT.restart.<rank>.nc (rank = PE_YIND NXPE + PE_XIND): ".<rank>" goes in front of the extension, or replaces a
  // ".0" already there
  auto rankPath = [](std::string path) {
    Mesh* gm = bout::globals::mesh;
    if (path.empty() || gm->NYPE * gm->NXPE == 1) {
      return path;
    }
    const int rank = gm->PE_YIND * gm->NXPE + gm->PE_XIND;
    const size_t dot = path.rfind('.');
    std::string stem = dot == std::string::npos ? path : path.substr(0, dot);
    const std::string ext = dot == std::string::npos ? std::string() : path.substr(dot);
    if (stem.size() >= 2 && stem.compare(stem.size() - 2, 2, ".0") == 0) {
      stem.erase(stem.size() - 2);
    }
    return stem + "." + std::to_string(rank) + ext;
  };
  const std::string restart_path = rankPath(ro.isSet("path") ? ro["path"].as<std::string>() : std::string());
  int iteration0 = 0;
  if (Options::root()["restart"].withDefault(false)) {
    if (restart_path.empty()) {
      throw BoutException("restart = true needs [restart_file] path");
    }
    const auto vars = bout::b200::readNcFixedVariables(restart_path);
    for (auto& f : f3d) {
      const auto it = vars.find(f.name);
      if (it == vars.end() || it->second.data.size() != f.var->size()) {
        throw BoutException("Restart file '{}' has no variable '{}' of the mesh's size", restart_path, f.name);
      }
      f.var->setHost(it->second.data);
    }
    if (vars.count("tt") == 0 || vars.count("hist_hi") == 0) {
      throw BoutException("Restart file '{}' lacks tt / hist_hi", restart_path);
    }
    simtime = vars.at("tt").data.at(0);
    iteration0 = static_cast<int>(vars.at("hist_hi").data.at(0)) + 1;
  }
  init_solver();
  if (run_rhs(simtime) != 0) { // solver.cxx:569
    throw BoutException("Physics RHS call failed\n");
  }
  // Output step ([output] path = <file>): evolving variables (Solver::outputVars, solver.cxx:694-724) and the fields
  // the model registers in outputVars(), one record at the start (call_monitors(simtime, -1, NOUT), solver.cxx:580)
  // and one per output step; device -> host -> disk overlaps the following steps (src/dump_writer.cxx).
  std::unique_ptr<bout::b200::OutputPipeline> dump;
  std::vector<const Field3D*> dump_fields;
  Options& oo = Options::root()["output"];
  const std::string dump_path = rankPath(oo.isSet("path") ? oo["path"].as<std::string>() : std::string());
  if (!dump_path.empty() || !restart_path.empty()) {
    std::vector<std::string> names;
    for (const auto& f : f3d) {
      names.push_back(f.name);
      dump_fields.push_back(f.var);
    }
    Options registered;
    model->outputVars(registered);
    for (auto& kv : registered.sections()) {
      const bool evolving = std::find(names.begin(), names.end(), kv.first) != names.end();
      if (kv.second.field3d != nullptr && !evolving) {
        names.push_back(kv.first);
        dump_fields.push_back(kv.second.field3d);
      }
    }
    const DataFileFacade& reg = model->dump; // SAVE_REPEAT / SAVE_ONCE registrations
    if (!reg.unsupported.empty()) {
      throw BoutException("Output step: SAVE_REPEAT of scalar / Field2D variable '{}' is not supported", reg.unsupported[0]);
    }
    for (const auto& kv : reg.repeat3d) {
      if (std::find(names.begin(), names.end(), kv.first) == names.end()) {
        names.push_back(kv.first);
        dump_fields.push_back(kv.second);
      }
    }
    std::vector<bout::b200::DumpScalar> scalars;
    std::vector<bout::b200::DumpOnceField> once;
    for (const auto& kv : reg.once_int) {
      scalars.push_back({kv.first, false, *kv.second, 0.0});
    }
    for (const auto& kv : reg.once_real) {
      scalars.push_back({kv.first, true, 0, *kv.second});
    }
    for (const auto& kv : reg.once2d) {
      if (kv.second->isAllocated()) {
        once.push_back({kv.first, false, kv.second->getHost()});
      }
    }
    for (const auto& kv : reg.once3d) {
      if (kv.second->isAllocated()) {
        once.push_back({kv.first, true, kv.second->getHost()});
      }
    }
    dump = std::make_unique<bout::b200::OutputPipeline>(dump_path, f3d[0].var->getMesh(), names, scalars, once, restart_path,
                                                        f3d.size());
    dump->snapshot(simtime, iteration0 - 1, dump_fields);
  }
  call_monitors(simtime, -1, nout); // solver.cxx:580
  for (int s = 0; s < nout; ++s) {
    run_output_step(simtime + output_timestep);
    run_rhs(simtime); // rk4.cxx:157: extra variables at the output time
    if (dump) {
      dump->snapshot(simtime, iteration0 + s, dump_fields);
    }
    call_monitors(simtime, s, nout);
  }
  if (dump) {
    dump->finish();
  }
  check(bout_b200_stream_sync(nullptr));
  return 0;
}

namespace {
void rkStage(Mesh* m, int mode, double dt, double c, const Field3D& start, const Field3D* k1, const Field3D* k2,
             const Field3D* k3, const Field3D* k4, Field3D& out) {
  const auto g = m->grid();
  check(bout_b200_rk_stage(&g, BOUT_B200_RGN_NOBNDRY, mode, dt, c, start.dptr(), k1 ? k1->dptr() : nullptr,
                           k2 ? k2->dptr() : nullptr, k3 ? k3->dptr() : nullptr, k4 ? k4->dptr() : nullptr, out.wptr(),
                           nullptr));
}

// RK4 with the stage updates carried by the right-hand side itself (StageFusion): stage j reads the state in the
// model's own variables, and its kernel writes the next stage state (start + a_j dt k_j) and the running result
// (start + sum b_i dt k_i) -- no k arrays, no axpy passes, no final combination (rk4.cxx:172-210 accumulated stage by
// stage, as bout_b200_rk_epilogue documents).  Stage states move between the variables and two work fields as block
// handles.  Returns false, having evaluated the first right-hand side, when the model's kernel did not take the stage.
bool rk4StepFused(Solver& s, BoutReal t, BoutReal dt) {
  auto& f3d = s.f3d;
  const size_t nv = f3d.size();
  Mesh* m = f3d[0].var->getMesh();
  bout::b200::StageFusion& sf = bout::b200::stageFusion();
  std::vector<Field3D> y0(nv), sa(nv), sb(nv), acc(nv);
  for (size_t i = 0; i < nv; ++i) {
    sa[i] = Field3D(m);
    sa[i].allocate();
    acc[i] = Field3D(m);
    acc[i].allocate();
  }
  auto ptrs = [&](std::vector<Field3D>& v) {
    std::vector<Field3D*> p(nv);
    for (size_t i = 0; i < nv; ++i) {
      p[i] = &v[i];
    }
    return p;
  };
  const std::vector<Field3D*> none(nv, nullptr);
  std::vector<Field3D*> vars(nv);
  for (size_t i = 0; i < nv; ++i) {
    vars[i] = f3d[i].var;
  }
  // one stage: returns whether the model's kernel applied the update; if not, it is applied here from ddt
  auto stage = [&](BoutReal tt, const std::vector<Field3D*>& start, const std::vector<Field3D*>& acc_in,
                   const std::vector<Field3D*>& state_out, const std::vector<Field3D*>& acc_out, double a, double b,
                   bool may_fall_back) {
    sf = bout::b200::StageFusion{};
    sf.active = true;
    sf.vars = vars;
    sf.start = start;
    sf.acc_in = acc_in;
    sf.state_out = state_out;
    sf.acc_out = acc_out;
    sf.a = a;
    sf.b = b;
    try {
      s.run_rhs(tt);
    } catch (...) {
      sf = bout::b200::StageFusion{};
      throw;
    }
    const bool consumed = sf.consumed;
    if (consumed) { // ddt() was not written by the kernel: nothing may have been added to it afterwards
      for (size_t e = 0; e < sf.ddt_blocks.size(); ++e) {
        const Field3D* d = sf.vars[e]->timeDeriv();
        if (d->block() != sf.ddt_blocks[e] || d->block()->version != sf.ddt_versions[e]) {
          sf = bout::b200::StageFusion{};
          throw BoutException("rk4: the model changed ddt() after its fused right-hand-side loop, which the fused stage "
                              "update cannot see; run with BOUT_B200_RK_FUSE=0");
        }
      }
    }
    sf = bout::b200::StageFusion{};
    if (consumed || may_fall_back) {
      return consumed;
    }
    for (size_t i = 0; i < nv; ++i) { // the same update from ddt, variable by variable
      const Field3D k = *f3d[i].F_var;
      const Field3D& base = acc_in[i] != nullptr ? *acc_in[i] : *start[i];
      if (state_out[i] != nullptr) {
        rkStage(m, 0, a, 1.0, *start[i], &k, nullptr, nullptr, nullptr, *state_out[i]);
      }
      if (acc_out[i] != nullptr) {
        if (acc_out[i] == acc_in[i] || acc_out[i] == start[i]) { // in place: through a private handle on the operand
          const Field3D keep = base;
          rkStage(m, 0, b, 1.0, keep, &k, nullptr, nullptr, nullptr, *acc_out[i]);
        } else {
          rkStage(m, 0, b, 1.0, base, &k, nullptr, nullptr, nullptr, *acc_out[i]);
        }
      }
    }
    return false;
  };
  // stage 1 reads the variables themselves as `start` (boundary conditions / exchanges only touch their guard cells)
  if (!stage(t, vars, none, ptrs(sa), ptrs(acc), 0.5 * dt, dt / 6.0, s.stage_fusion < 0) && s.stage_fusion < 0) {
    return false;
  }
  for (size_t i = 0; i < nv; ++i) {
    y0[i] = std::move(*f3d[i].var);
    *f3d[i].var = std::move(sa[i]);
    sb[i] = Field3D(m);
    sb[i].allocate();
  }
  stage(t + 0.5 * dt, ptrs(y0), ptrs(acc), ptrs(sb), ptrs(acc), 0.5 * dt, dt / 3.0, false);
  for (size_t i = 0; i < nv; ++i) {
    sa[i] = std::move(*f3d[i].var);
    *f3d[i].var = std::move(sb[i]);
  }
  stage(t + 0.5 * dt, ptrs(y0), ptrs(acc), ptrs(sa), ptrs(acc), dt, dt / 3.0, false);
  for (size_t i = 0; i < nv; ++i) {
    sb[i] = std::move(*f3d[i].var);
    *f3d[i].var = std::move(sa[i]);
  }
  stage(t + dt, ptrs(y0), ptrs(acc), none, ptrs(y0), 0.0, dt / 6.0, false); // the result lands in the start block
  for (size_t i = 0; i < nv; ++i) {
    *f3d[i].var = std::move(y0[i]);
  }
  return true;
}

void rk4Step(Solver& s, BoutReal t, BoutReal dt) { // rk4.cxx:172-210 field by field
  auto& f3d = s.f3d;
  const size_t nv = f3d.size();
  std::vector<Field3D> start(nv), k1(nv), k2(nv), k3(nv), k4(nv);
  Mesh* m = f3d[0].var->getMesh();
  static const bool fuse_env = [] {
    const char* e = getenv("BOUT_B200_RK_FUSE");
    return e == nullptr || e[0] != '0';
  }();
  bool have_k1 = false;
  if (fuse_env && nv == 2 && s.stage_fusion != 0) {
    if (rk4StepFused(s, t, dt)) {
      s.stage_fusion = 1;
      return;
    }
    s.stage_fusion = 0; // the first right-hand side was evaluated the ordinary way: carry on from its ddt
    have_k1 = true;
  }
  for (size_t i = 0; i < nv; ++i) {
    start[i] = *f3d[i].var;
    start[i].allocate(); // a private copy: the variable is overwritten in place below
  }
  auto stage = [&](std::vector<Field3D>& k, BoutReal tt) {
    if (!(have_k1 && &k == &k1)) {
      s.run_rhs(tt);
    }
    for (size_t i = 0; i < nv; ++i) {
      k[i] = *f3d[i].F_var;
    }
  };
  auto advance = [&](const std::vector<Field3D>& k, BoutReal c) { // var(interior) = start + c dt k; guards kept
    for (size_t i = 0; i < nv; ++i) {
      rkStage(m, 0, dt, c, start[i], &k[i], nullptr, nullptr, nullptr, *f3d[i].var);
    }
  };
  stage(k1, t);
  advance(k1, 0.5);
  stage(k2, t + 0.5 * dt);
  advance(k2, 0.5);
  stage(k3, t + 0.5 * dt);
  advance(k3, 1.0);
  stage(k4, t + dt);
  for (size_t i = 0; i < nv; ++i) {
    rkStage(m, 1, dt, 1.0, start[i], &k1[i], &k2[i], &k3[i], &k4[i], *f3d[i].var);
  }
}

class RK4Solver : public Solver {
public:
  using Solver::Solver;
  void take_step(BoutReal t, BoutReal dt) override { rk4Step(*this, t, dt); }
};

class RK3SSP : public Solver {
public:
  using Solver::Solver;
  void take_step(BoutReal t, BoutReal dt) override { // rk3-ssp.cxx:100-129
    const size_t nv = f3d.size();
    std::vector<Field3D> start(nv), u(nv);
    Mesh* m = f3d[0].var->getMesh();
    for (size_t i = 0; i < nv; ++i) {
      start[i] = *f3d[i].var;
      start[i].allocate();
    }
    run_rhs(t);
    for (size_t i = 0; i < nv; ++i) { // u1 = start + dt L
      rkStage(m, 0, dt, 1.0, start[i], f3d[i].F_var, nullptr, nullptr, nullptr, *f3d[i].var);
    }
    run_rhs(t + dt);
    for (size_t i = 0; i < nv; ++i) { // u2 = 3/4 start + 1/4 u1 + 1/4 dt L
      u[i] = *f3d[i].var;
      u[i].allocate();
      rkStage(m, 2, dt, 1.0, start[i], &u[i], f3d[i].F_var, nullptr, nullptr, *f3d[i].var);
    }
    run_rhs(t + 0.5 * dt);
    for (size_t i = 0; i < nv; ++i) { // result = 1/3 start + 2/3 (u2 + dt L)
      u[i] = *f3d[i].var;
      u[i].allocate();
      rkStage(m, 3, dt, 1.0, start[i], &u[i], f3d[i].F_var, nullptr, nullptr, *f3d[i].var);
    }
  }
};

// ------------------------------------------------------------------------------------------------------
// Other explicit integrators: euler, rkgeneric (rkf45 / cashkarp / rk4 / rkf34), splitrk.
// A "state" is one Field3D handle per evolving variable.  New states are always formed INTO the evolving
// variables themselves (bout_b200_lincomb over RGN_NOBNDRY through the copy-on-write handle, so guard cells carry
// over exactly as with the reference's load_vars, solver.cxx:1039-1140, which writes the evolving region only);
// keeping or re-loading a state is a handle copy, not a pass over memory.
// ------------------------------------------------------------------------------------------------------
using State = std::vector<Field3D>;
struct Term {
  double c;
  const Field3D* f;
};

// out(RGN_NOBNDRY) = sum_j c_j f_j, terms added left to right; operands may share storage with `out`
void combine(Field3D& out, const std::vector<Term>& terms) {
  Mesh* m = out.getMesh();
  const auto g = m->grid();
  std::vector<double> c;
  std::vector<const double*> x;
  for (const auto& t : terms) {
    if (t.c == 0.0) {
      continue; // rkscheme.cxx:179 skips zero coefficients; also keeps never-written stage fields out
    }
    if (!t.f->isAllocated()) {
      throw BoutException("Solver: state field used before it was set");
    }
    c.push_back(t.c);
    x.push_back(t.f->dptr()); // taken BEFORE un-sharing `out`: an operand that shares out's block keeps reading the old block
  }
  if (c.empty()) {
    throw BoutException("Solver: empty linear combination");
  }
  Field3D keep;
  for (const double* xp : x) {
    if (xp == out.dptr()) {
      keep = out; // an operand IS out's block: keep it alive and force the un-share below
    }
  }
  double* w = out.wptr(); // fresh block with the old contents (guards included) whenever the block was shared
  constexpr size_t kMax = 8; // bout_b200_lincomb takes up to 8 terms; longer sums continue from the partial result
  if (c.size() <= kMax) {
    check(bout_b200_lincomb(&g, BOUT_B200_RGN_NOBNDRY, static_cast<int>(c.size()), c.data(), x.data(), w, nullptr));
    return;
  }
  Field3D partial = emptyFrom(out), next;
  size_t done = 0;
  check(bout_b200_lincomb(&g, BOUT_B200_RGN_NOBNDRY, static_cast<int>(kMax), c.data(), x.data(), partial.wptr(), nullptr));
  done = kMax;
  while (done < c.size()) {
    const size_t n = std::min(kMax - 1, c.size() - done);
    std::vector<double> cc{1.0};
    std::vector<const double*> xx{partial.dptr()};
    cc.insert(cc.end(), c.begin() + done, c.begin() + done + n);
    xx.insert(xx.end(), x.begin() + done, x.begin() + done + n);
    done += n;
    if (done == c.size()) {
      check(bout_b200_lincomb(&g, BOUT_B200_RGN_NOBNDRY, static_cast<int>(cc.size()), cc.data(), xx.data(), w, nullptr));
    } else {
      next = emptyFrom(out);
      check(bout_b200_lincomb(&g, BOUT_B200_RGN_NOBNDRY, static_cast<int>(cc.size()), cc.data(), xx.data(), next.wptr(), nullptr));
      partial = next;
    }
  }
}

// Local error sum of the adaptive integrators on the device, one number back to the host
class ErrorNorm {
public:
  // sum_i sum_{RGN_NOBNDRY} |a_i - b_i| / (|a_i| + |b_i| + atol) / neq  (rkscheme.cxx:139-168, split-rk.cxx:127-144)
  double operator()(const State& a, const State& b, double atol) {
    Mesh* m = a.at(0).getMesh();
    const auto g = m->grid();
    if (!scratch) {
      scratch = bout::b200::acquire(bout_b200_err_sum_scratch_doubles() + 2);
    }
    double* result = scratch->ptr + bout_b200_err_sum_scratch_doubles();
    for (size_t i = 0; i < a.size(); ++i) {
      check(bout_b200_err_sum(&g, BOUT_B200_RGN_NOBNDRY, a[i].dptr(), b[i].dptr(), atol, scratch->ptr, i == 0 ? 0 : 1,
                              result, nullptr));
    }
    double host = 0.0;
    check(bout_b200_memcpy_d2h(&host, result, sizeof(double), nullptr));
    check(bout_b200_stream_sync(nullptr));
    double neq = static_cast<double>(a.size()) * (g.xend - g.xstart + 1) * (g.yend - g.ystart + 1) * g.nz;
    if (m->NYPE * m->NXPE > 1) { // the reference sums error and neq over the ranks (rkscheme.cxx:155-163, MPI_Allreduce)
      host = globalReduce(host, 0);
      neq *= m->NYPE * m->NXPE;
    }
    return host / neq;
  }

  // max_i max_{RGN_NOBNDRY} |a_i - b_i|  (adams_bashforth.cxx:229-253)
  double maxAbsDiff(const State& a, const State& b) {
    Mesh* m = a.at(0).getMesh();
    const auto g = m->grid();
    if (!scratch) {
      scratch = bout::b200::acquire(bout_b200_err_sum_scratch_doubles() + 2);
    }
    double* result = scratch->ptr + bout_b200_err_sum_scratch_doubles();
    for (size_t i = 0; i < a.size(); ++i) {
      check(bout_b200_err_max(&g, BOUT_B200_RGN_NOBNDRY, a[i].dptr(), b[i].dptr(), scratch->ptr, i == 0 ? 0 : 1, result,
                              nullptr));
    }
    double host = 0.0;
    check(bout_b200_memcpy_d2h(&host, result, sizeof(double), nullptr));
    check(bout_b200_stream_sync(nullptr));
    return m->NYPE * m->NXPE > 1 ? globalReduce(host, 1) : host;
  }

  static double globalReduce(double v, int op) {
    const bout::b200::YDecomposition& yd = bout::b200::yDecomposition();
    if (yd.allreduce == nullptr) {
      throw BoutException("adaptive time stepping on {} ranks needs YDecomposition::allreduce", yd.ranks());
    }
    return yd.allreduce(v, op, yd.user);
  }

private:
  std::shared_ptr<bout::b200::Block> scratch;
};

class StateSolver : public Solver {
public:
  using Solver::Solver;

protected:
  size_t nv() const { return f3d.size(); }
  State current() const { // handles of the evolving variables (no copy)
    State s(nv());
    for (size_t i = 0; i < nv(); ++i) {
      s[i] = *f3d[i].var;
    }
    return s;
  }
  State derivs() const {
    State s(nv());
    for (size_t i = 0; i < nv(); ++i) {
      s[i] = *f3d[i].F_var;
    }
    return s;
  }
  void load(const State& s) { // load_vars
    for (size_t i = 0; i < nv(); ++i) {
      if (f3d[i].var->dptr() != s[i].dptr()) {
        *f3d[i].var = s[i];
      }
    }
  }
  /// variables <- sum_j c_j states_j (per variable)
  void form(const std::vector<std::pair<double, const State*>>& terms) {
    for (size_t i = 0; i < nv(); ++i) {
      std::vector<Term> t;
      for (const auto& p : terms) {
        t.push_back({p.first, &(*p.second)[i]});
      }
      combine(*f3d[i].var, t);
    }
  }
};

// ---- euler (src/solver/impls/euler/euler.cxx) ---------------------------------------------------------------
class EulerSolver : public StateSolver {
public:
  explicit EulerSolver(Options* opts) : StateSolver(opts) {
    mxstep = (*options)["mxstep"].withDefault(500);
    cfl_factor = (*options)["cfl_factor"].withDefault(2.);
  }
  void setMaxTimestep(BoutReal dt) override { // euler.cxx:28-36
    if (dt >= cfl_factor * timestep) {
      return;
    }
    timestep = dt * 0.99 / cfl_factor;
    timestep_reduced = true;
  }
  void run_output_step(BoutReal target) override { // euler.cxx:67-135
    bool running = true;
    int internal_steps = 0;
    do {
      BoutReal dt_limit = timestep;
      if ((simtime + timestep) >= target) {
        timestep = target - simtime;
        running = false;
      }
      const BoutReal old_timestep = timestep;
      const State start = current();
      timestep_reduced = false;
      eulerStep(simtime, timestep, start);
      if (timestep_reduced && timestep < old_timestep) { // the model lowered the limit during the RHS: redo
        eulerStep(simtime, timestep, start);
        dt_limit = timestep;
        running = true;
      }
      simtime += timestep;
      if (++internal_steps > mxstep) {
        throw BoutException("ERROR: MXSTEP exceeded. simtime={:e}, timestep = {:e}\n", simtime, timestep);
      }
      timestep = dt_limit;
    } while (running);
  }

private:
  int mxstep;
  BoutReal cfl_factor;
  bool timestep_reduced = false;
  void eulerStep(BoutReal t, BoutReal dt, const State& start) { // euler.cxx:138-182: result = start + dt * F(start)
    load(start);
    run_rhs(t);
    const State k = derivs();
    form({{1.0, &start}, {dt, &k}});
  }
};

// ---- rkgeneric (src/solver/impls/rkgeneric/rkgeneric.cxx, rkscheme.cxx and impls/*) ------------------------------
struct RKTableau {
  std::string label;
  int stages = 0, order = 0;
  std::vector<double> stage;  // [stages][stages]
  std::vector<double> result; // [stages][2]: column 0 / 1 as in RKScheme::resultCoeffs
  std::vector<double> time;   // [stages]
  double a(int i, int j) const { return stage[static_cast<size_t>(i) * stages + j]; }
  double b(int i, int k) const { return result[static_cast<size_t>(i) * 2 + k]; }
};

RKTableau makeTableau(const std::string& name) {
  RKTableau t;
  auto size = [&](int n, int order) {
    t.stages = n;
    t.order = order;
    t.stage.assign(static_cast<size_t>(n) * n, 0.0);
    t.result.assign(static_cast<size_t>(n) * 2, 0.0);
    t.time.assign(n, 0.0);
  };
  auto A = [&](int i, int j) -> double& { return t.stage[static_cast<size_t>(i) * t.stages + j]; };
  auto B = [&](int i, int k) -> double& { return t.result[static_cast<size_t>(i) * 2 + k]; };
  t.label = name;
  if (name == "rkf45") { // Fehlberg 4(5): impls/rkf45/rkf45.cxx:27-88
    size(6, 4);
    A(1, 0) = 1.0 / 4.0;
    A(2, 0) = 3.0 / 32.0, A(2, 1) = 9.0 / 32.0;
    A(3, 0) = 1932.0 / 2197.0, A(3, 1) = -7200.0 / 2197.0, A(3, 2) = 7296.0 / 2197.0;
    A(4, 0) = 439.0 / 216.0, A(4, 1) = -8.0, A(4, 2) = 3680.0 / 513.0, A(4, 3) = -845.0 / 4104.0;
    A(5, 0) = -8.0 / 27.0, A(5, 1) = 2.0, A(5, 2) = -3544.0 / 2565.0, A(5, 3) = 1859.0 / 4104.0, A(5, 4) = -11.0 / 40.0;
    B(0, 0) = 16.0 / 135.0, B(0, 1) = 25.0 / 216.0;
    B(2, 0) = 6656.0 / 12825.0, B(2, 1) = 1408.0 / 2565.0;
    B(3, 0) = 28561.0 / 56430.0, B(3, 1) = 2197.0 / 4104.0;
    B(4, 0) = -9.0 / 50.0, B(4, 1) = -1.0 / 5.0;
    B(5, 0) = 2.0 / 55.0;
    t.time = {0.0, 1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
  } else if (name == "cashkarp") { // impls/cashkarp/cashkarp.cxx:27-88
    size(6, 4);
    A(1, 0) = 1.0 / 5.0;
    A(2, 0) = 3.0 / 40.0, A(2, 1) = 9.0 / 40.0;
    A(3, 0) = 3.0 / 10.0, A(3, 1) = -9.0 / 10.0, A(3, 2) = 6.0 / 5.0;
    A(4, 0) = -11.0 / 54.0, A(4, 1) = 5.0 / 2.0, A(4, 2) = -70.0 / 27.0, A(4, 3) = 35.0 / 27.0;
    A(5, 0) = 1631.0 / 55296.0, A(5, 1) = 175.0 / 512.0, A(5, 2) = 575.0 / 13824.0, A(5, 3) = 44275.0 / 110592.0,
         A(5, 4) = 253.0 / 4096.0;
    B(0, 0) = 37.0 / 378.0, B(0, 1) = 2825.0 / 27648.0;
    B(2, 0) = 250.0 / 621.0, B(2, 1) = 18575.0 / 48384.0;
    B(3, 0) = 125.0 / 594.0, B(3, 1) = 13525.0 / 55296.0;
    B(4, 1) = 277.0 / 14336.0;
    B(5, 0) = 512.0 / 1771.0, B(5, 1) = 1.0 / 4.0;
    t.time = {0.0, 1.0 / 5.0, 3.0 / 10.0, 3.0 / 5.0, 1.0, 7.0 / 8.0};
  } else if (name == "rkf34") { // impls/rkf34/rkf34.cxx:33-83
    size(5, 3);
    A(1, 0) = 1.0 / 4.0;
    A(2, 0) = 4.0 / 81.0, A(2, 1) = 32.0 / 81.0;
    A(3, 0) = 57.0 / 98.0, A(3, 1) = -432.0 / 343.0, A(3, 2) = 1053.0 / 686.0;
    A(4, 0) = 1.0 / 6.0, A(4, 2) = 27.0 / 52.0, A(4, 3) = 49.0 / 156.0;
    B(0, 0) = 43.0 / 288.0, B(0, 1) = 1.0 / 6.0;
    B(2, 0) = 243.0 / 416.0, B(2, 1) = 27.0 / 52.0;
    B(3, 0) = 343.0 / 1872.0, B(3, 1) = 49.0 / 156.0;
    B(4, 0) = 1.0 / 12.0;
    t.time = {0.0, 1.0 / 4.0, 4.0 / 9.0, 6.0 / 7.0, 1.0};
  } else if (name == "rk4") { // one RK4 step against two half steps: impls/rk4simple/rk4simple.cxx:27-116
    size(11, 4);
    A(1, 0) = 1.0 / 2.0, A(2, 1) = 1.0 / 2.0, A(3, 2) = 1.0;
    A(4, 0) = 1.0 / 4.0, A(5, 4) = 1.0 / 4.0, A(6, 5) = 1.0 / 2.0;
    for (int i = 7; i <= 10; ++i) { // the state after the first half step
      A(i, 0) = 1.0 / 12.0, A(i, 4) = 1.0 / 6.0, A(i, 5) = 1.0 / 6.0, A(i, 6) = 1.0 / 12.0;
    }
    A(8, 7) = 1.0 / 4.0, A(9, 8) = 1.0 / 4.0, A(10, 9) = 1.0 / 2.0;
    B(0, 1) = 1.0 / 6.0, B(1, 1) = 1.0 / 3.0, B(2, 1) = 1.0 / 3.0, B(3, 1) = 1.0 / 6.0;
    B(0, 0) = 1.0 / 12.0, B(4, 0) = 1.0 / 6.0, B(5, 0) = 1.0 / 6.0, B(6, 0) = 1.0 / 12.0;
    B(7, 0) = 1.0 / 12.0, B(8, 0) = 1.0 / 6.0, B(9, 0) = 1.0 / 6.0, B(10, 0) = 1.0 / 12.0;
    t.time = {0.0, 0.5, 0.5, 1.0, 0.25, 0.25, 0.5, 0.5, 0.75, 0.75, 1.0};
  } else {
    throw BoutException("RKScheme '{}' is not available: rkf45, cashkarp, rk4, rkf34", name);
  }
  return t;
}

class RKGenericSolver : public StateSolver {
public:
  explicit RKGenericSolver(Options* opts) : StateSolver(opts) { // rkgeneric.cxx:9-24, rkscheme.cxx:19-25
    atol = (*options)["atol"].withDefault(1.e-5);
    rtol = (*options)["rtol"].withDefault(1.e-3);
    max_timestep = (*options)["max_timestep"].withDefault(output_timestep);
    timestep = (*options)["timestep"].withDefault(max_timestep);
    mxstep = (*options)["mxstep"].withDefault(500);
    adaptive = (*options)["adaptive"].withDefault(true);
    tab = makeTableau((*options)["scheme"].withDefault(std::string("rkf45")));
    followHighOrder = (*options)["followHighOrder"].withDefault(false);
    dtfac = (*options)["dtfac"].withDefault(1.0);
    if (tab.label == "rkf34") {
      dtfac = 0.9; // impls/rkf34/rkf34.cxx:10-14 overrides the option
    }
  }
  void setMaxTimestep(BoutReal dt) override { // rkgeneric.cxx:26-34
    if (dt > timestep) {
      return;
    }
    if (adaptive) {
      timestep = dt;
    }
  }
  void run_output_step(BoutReal target) override { // rkgeneric.cxx:82-160
    bool running = true;
    int internal_steps = 0;
    do {
      BoutReal dt;
      const State start = current();
      do {
        dt = timestep;
        running = true;
        if ((simtime + dt) >= target) {
          dt = target - simtime;
          running = false;
        }
        const BoutReal err = rkStep(simtime, dt, start);
        if (!adaptive) {
          break;
        }
        if (++internal_steps > mxstep) {
          throw BoutException("ERROR: MXSTEP exceeded. timestep = {:e}, err={:e}\n", timestep, err);
        }
        if ((err > rtol) || ((err < 0.1 * rtol) && running)) {
          timestep = dtfac * dt * std::pow(rtol / (2.0 * err), 1.0 / (tab.order + 1.0)); // rkscheme.cxx:129-131
          if ((max_timestep > 0) && (timestep > max_timestep)) {
            timestep = max_timestep;
          }
        }
        if (err < rtol) {
          break;
        }
      } while (true);
      simtime += dt; // the accepted result already sits in the evolving variables
    } while (running);
  }

private:
  BoutReal atol, rtol, max_timestep, dtfac;
  int mxstep;
  bool adaptive, followHighOrder;
  RKTableau tab;
  ErrorNorm errnorm;

  // rkgeneric.cxx:163-180 + rkscheme.cxx:52-127: stages, then the followed result into the variables
  BoutReal rkStep(BoutReal t, BoutReal dt, const State& start) {
    std::vector<State> k(tab.stages);
    for (int s = 0; s < tab.stages; ++s) {
      std::vector<std::pair<double, const State*>> terms{{1.0, &start}};
      for (int j = 0; j < s; ++j) {
        if (std::abs(tab.a(s, j)) < atol) { // rkscheme.cxx:74 (sic: compared with atol)
          continue;
        }
        terms.push_back({tab.a(s, j) * dt, &k[j]});
      }
      if (terms.size() == 1) {
        load(start);
      } else {
        form(terms);
      }
      run_rhs(t + dt * tab.time[s]);
      k[s] = derivs();
    }
    const int follow = followHighOrder ? 0 : 1, alt = followHighOrder ? 1 : 0;
    auto output = [&](int col) {
      std::vector<std::pair<double, const State*>> terms{{1.0, &start}};
      for (int s = 0; s < tab.stages; ++s) {
        terms.push_back({dt * tab.b(s, col), &k[s]}); // zero coefficients are dropped by combine()
      }
      form(terms);
    };
    if (!adaptive) {
      output(follow);
      return 0.0;
    }
    output(alt);
    const State resultAlt = current();
    output(follow);
    return errnorm(current(), resultAlt, atol);
  }
};

// ---- splitrk (src/solver/impls/split-rk/split-rk.cxx): Strang split, RK-Legendre diffusive + SSP-RK3 advective ---
class SplitRK : public StateSolver {
public:
  explicit SplitRK(Options* opts) : StateSolver(opts) { // split-rk.cxx:18-52
    nstages = (*options)["nstages"].withDefault(10);
    adaptive = (*options)["adaptive"].withDefault(true);
    atol = (*options)["atol"].withDefault(1e-10);
    rtol = (*options)["rtol"].withDefault(1e-5);
    max_timestep = (*options)["max_timestep"].withDefault(output_timestep);
    max_timestep_change = (*options)["max_timestep_change"].withDefault(2.0);
    mxstep = (*options)["mxstep"].withDefault(1000);
    adapt_period = (*options)["adapt_period"].withDefault(1);
    if (!(max_timestep_change > 1.0) || mxstep <= 0 || nstages <= 1 || adapt_period <= 0) {
      throw BoutException("splitrk: need max_timestep_change > 1, mxstep > 0, nstages > 1, adapt_period > 0");
    }
  }
  void init_solver() override { // split-rk.cxx:90-94
    const int ninternal_steps = static_cast<int>(std::ceil(output_timestep / timestep));
    if (ninternal_steps <= 0) {
      throw BoutException("splitrk: output step / timestep must be positive");
    }
    timestep = output_timestep / ninternal_steps;
  }
  void run_output_step(BoutReal target) override { // split-rk.cxx:99-207
    BoutReal dt;
    bool running = true;
    int internal_steps = 0;
    do {
      if (adaptive && (internal_steps % adapt_period == 0)) {
        const State state = current();
        do {
          dt = timestep;
          running = true;
          if ((simtime + dt) >= target) {
            dt = target - simtime;
            running = false;
          }
          load(state); // two half steps
          splitStep(simtime, 0.5 * dt);
          splitStep(simtime + 0.5 * dt, 0.5 * dt);
          const State state2 = current();
          load(state); // one full step
          splitStep(simtime, dt);
          const State state1 = current();
          const BoutReal err = errnorm(state2, state1, atol);
          if (++internal_steps > mxstep) {
            throw BoutException("ERROR: MXSTEP exceeded. timestep = {:e}, err={:e}\n", timestep, err);
          }
          if ((err > rtol) || (err < 0.1 * rtol)) { // error ~ dt^2
            BoutReal factor = std::pow((0.5 * rtol) / err, 1. / 3);
            if (factor > max_timestep_change) {
              factor = max_timestep_change;
            } else if (factor < 1. / max_timestep_change) {
              factor = 1. / max_timestep_change;
            }
            timestep *= factor;
            if ((max_timestep > 0) && (timestep > max_timestep)) {
              timestep = max_timestep;
            }
          }
          if (err < rtol) {
            load(state2);
            break;
          }
        } while (true);
      } else {
        dt = timestep;
        running = true;
        if ((simtime + dt) >= target) {
          dt = target - simtime;
          running = false;
        }
        splitStep(simtime, timestep); // split-rk.cxx:192 (sic: the un-clipped timestep)
        internal_steps++;
      }
      simtime += dt;
    } while (running);
  }

private:
  int nstages, mxstep, adapt_period;
  bool adaptive;
  BoutReal atol, rtol, max_timestep, max_timestep_change;
  ErrorNorm errnorm;

  void splitStep(BoutReal t, BoutReal dt) { // split-rk.cxx:209-219, on the state held by the variables
    diffusionStep(t, 0.5 * dt);
    advectionStep(t, dt);
    diffusionStep(t + 0.5 * dt, 0.5 * dt);
  }

  void diffusionStep(BoutReal t, BoutReal dt) { // Runge-Kutta-Legendre, split-rk.cxx:221-283
    const BoutReal weight = dt * 4. / (SQ(static_cast<BoutReal>(nstages)) + nstages - 2);
    const State start = current();
    run_diffusive(t);
    const State dydt = derivs();
    form({{1.0, &start}, {weight / 3.0, &dydt}}); // y_1
    State u2 = current();
    run_diffusive(t + (weight / 3.0) * dt);
    State f = derivs();
    // y_2 = 1.5 (u2 + weight f(u2)) - 0.5 start - weight dydt
    form({{1.5, &u2}, {1.5 * weight, &f}, {-0.5, &start}, {-weight, &dydt}});
    State u1 = current();
    BoutReal b_jm2 = 1. / 3, b_jm1 = 1. / 3;
    for (int j = 3; j <= nstages; ++j) {
      const BoutReal b_j = (SQ(static_cast<BoutReal>(j)) + j - 2.0) / (2. * j * (j + 1.));
      const BoutReal mu = (2. * j - 1.) / j * b_j / b_jm1;
      const BoutReal nu = -(j - 1.) / j * b_j / b_jm2;
      const BoutReal a_jm1 = 1. - b_jm1;
      run_diffusive(t); // the variables hold u1
      f = derivs();
      // mu (u1 + weight (f - a_jm1 dydt)) + nu u2 + (1 - mu - nu) start
      form({{mu, &u1}, {mu * weight, &f}, {-mu * weight * a_jm1, &dydt}, {nu, &u2}, {1. - mu - nu, &start}});
      b_jm2 = b_jm1;
      b_jm1 = b_j;
      u2 = u1;
      u1 = current();
    }
  }

  void advectionStep(BoutReal t, BoutReal dt) { // SSP-RK3, split-rk.cxx:285-318
    const State start = current();
    run_convective(t);
    State k = derivs();
    form({{1.0, &start}, {dt, &k}});
    const State u1 = current();
    run_convective(t + dt);
    k = derivs();
    form({{0.75, &start}, {0.25, &u1}, {0.25 * dt, &k}});
    const State u2 = current();
    run_convective(t + 0.5 * dt);
    k = derivs();
    form({{1. / 3, &start}, {2. / 3., &u2}, {(2. / 3.) * dt, &k}});
  }
};

// ---- rk4 with solver:adaptive = true (rk4.cxx:84-140): two half steps against one full step ---------------------
class RK4Adaptive : public StateSolver {
public:
  explicit RK4Adaptive(Options* opts) : StateSolver(opts) { // rk4.cxx:11-25
    atol = (*options)["atol"].withDefault(1.e-5);
    rtol = (*options)["rtol"].withDefault(1.e-3);
    max_timestep = (*options)["max_timestep"].withDefault(output_timestep);
    timestep = (*options)["timestep"].withDefault(max_timestep);
    mxstep = (*options)["mxstep"].withDefault(500);
  }
  void setMaxTimestep(BoutReal dt) override { // rk4.cxx:27-35
    if (dt <= timestep) {
      timestep = dt;
    }
  }
  void run_output_step(BoutReal target) override {
    bool running = true;
    int internal_steps = 0;
    do {
      BoutReal dt;
      const State f0 = current();
      do {
        dt = timestep;
        running = true;
        if ((simtime + dt) >= target) {
          dt = target - simtime;
          running = false;
        }
        load(f0);
        rk4Step(*this, simtime, 0.5 * dt);
        rk4Step(*this, simtime + 0.5 * dt, 0.5 * dt);
        const State f2 = current();
        load(f0);
        rk4Step(*this, simtime, dt);
        const State f1 = current();
        const BoutReal err = errnorm(f2, f1, atol);
        if (++internal_steps > mxstep) {
          throw BoutException("ERROR: MXSTEP exceeded. timestep = {:e}, err={:e}\n", timestep, err);
        }
        if ((err > rtol) || (err < 0.1 * rtol)) { // error ~ dt^5
          timestep /= std::pow(err / (0.5 * rtol), 0.2);
          if ((max_timestep > 0) && (timestep > max_timestep)) {
            timestep = max_timestep;
          }
        }
        if (err < rtol) {
          load(f2);
          break;
        }
      } while (true);
      simtime += dt;
    } while (running);
  }

private:
  BoutReal atol, rtol, max_timestep;
  int mxstep;
  ErrorNorm errnorm;
};

// ---- adams-bashforth (src/solver/impls/adams_bashforth/adams_bashforth.cxx): variable-step, variable-order ------
// Coefficient of history point `position` for the advance from points[0] to `next`: the Lagrange basis polynomial
// through the first `order` time points, integrated with the closed Newton-Cotes rule on order + 1 nodes (exact for
// that degree) -- adams_bashforth.cxx:10-186.  Host scalar arithmetic only.
double abCoefficient(double next, const std::deque<double>& points, int position, int order) {
  static const std::vector<std::vector<double>> weights = {
      {1., 1.},
      {1., 4., 1.},
      {1., 3., 3., 1.},
      {7., 32., 12., 32., 7.},
      {19., 75., 50., 50., 75., 19.},
      {41., 216., 27., 272., 27., 216., 41.},
      {751., 3577., 1323., 2989., 2989., 1323., 3577., 751.},
      {989., 5888., -928., 10496., -4540., 10496., -928., 5888., 989.}};
  static const double scale[] = {1.0 / 2.0, 1.0 / 3.0, 3.0 / 8.0, 2.0 / 45.0, 5.0 / 288.0, 1.0 / 140.0, 7.0 / 17280.0,
                                 4.0 / 14175.0};
  const int rule = std::min(std::max(order, 1), 8) - 1; // order 1 -> trapezium ... order >= 8 -> 9-point rule
  const auto& w = weights[rule];
  const int npts = static_cast<int>(w.size()) - 1; // basis points entering the products (:149-166)
  const double start = points[0];
  const double step = (next - start) / (static_cast<double>(w.size()) - 1.0);
  double sum = 0.0;
  for (size_t i = 0; i < w.size(); ++i) {
    const double x = start + i * step;
    double num = 1.0;
    for (int k = 0; k < npts; ++k) {
      num *= (k != position) ? (x - points[k]) : 1.0;
    }
    sum += (w[i] * scale[rule]) * num;
  }
  double den = 1.0;
  for (int k = 0; k < npts; ++k) {
    den /= (k != position) ? (points[position] - points[k]) : 1.0;
  }
  return step * sum * den;
}

class AdamsBashforthSolver : public StateSolver {
public:
  explicit AdamsBashforthSolver(Options* opts) : StateSolver(opts) { // adams_bashforth.cxx:256-285
    atol = (*options)["atol"].withDefault(1.e-5);
    rtol = (*options)["rtol"].withDefault(1.e-5);
    mxstep = (*options)["mxstep"].withDefault(50000);
    adaptive = (*options)["adaptive"].withDefault(true);
    adaptive_order = (*options)["adaptive_order"].withDefault(true);
    followHighOrder = (*options)["followHighOrder"].withDefault(true);
    dtFac = (*options)["dtFac"].withDefault(0.75);
    maximum_order = (*options)["order"].withDefault(5);
    max_timestep = (*options)["max_timestep"].withDefault(output_timestep);
    timestep = (*options)["timestep"].withDefault(max_timestep);
    if (maximum_order < 1) {
      throw BoutException("adams-bashforth: order must be >= 1");
    }
  }
  void setMaxTimestep(BoutReal dt) override { // :289-300
    if (dt <= timestep && adaptive) {
      timestep = dt;
    }
  }
  void init_solver() override { // :320-336
    if (!adaptive && (output_timestep / timestep > mxstep)) {
      throw BoutException("ERROR: Requested timestep would lead to MXSTEP being exceeded. timestep = {:e}, MXSTEP={:d}\n",
                          timestep, mxstep);
    }
    current_order = 1;
  }
  void run_output_step(BoutReal target) override { // :352-540
    bool running = true;
    int internal_steps = 0;
    while (running) {
      run_rhs(simtime); // the derivative at the current time joins the history once per step
      history.push_front(derivs());
      times.push_front(simtime);
      const State state = current();
      State nextState;
      bool use_lower = false;
      BoutReal dt;
      while (true) {
        timestep = std::min(timestep, max_timestep);
        dt = timestep;
        running = true;
        if ((simtime + dt) >= target) {
          dt = target - simtime;
          running = false;
        }
        const BoutReal err = abStep(simtime, dt, current_order, state, nextState);
        if (!adaptive) {
          break;
        }
        if (++internal_steps > mxstep) {
          throw BoutException("ERROR: MXSTEP exceeded. timestep = {:e}, err={:e}\n", timestep, err);
        }
        BoutReal dt_lim = dt * timestepLimit(err, rtol, current_order);
        if (err < rtol) {
          if (adaptive_order && current_order > 1) { // would one order lower allow a longer step?
            State lowerNextState;
            const BoutReal lowerErr = abStep(simtime, dt, current_order - 1, state, lowerNextState);
            const BoutReal lower_dt_lim = dt * timestepLimit(lowerErr, rtol, current_order - 1);
            use_lower = lower_dt_lim > dt_lim;
            if (use_lower) {
              dt_lim = lower_dt_lim;
              nextState = lowerNextState;
              current_order = current_order - 1;
            }
          }
          timestep = std::min(timestep * 1.1, dt_lim);
          break;
        }
        timestep = 0.9 * dt_lim;
      }
      if (times.size() == static_cast<size_t>(maximum_order)) {
        times.pop_back();
      }
      if (history.size() == static_cast<size_t>(maximum_order)) {
        history.pop_back();
      }
      if (current_order < maximum_order && !use_lower) {
        current_order++;
      }
      simtime += dt;
      load(nextState);
    }
  }

private:
  BoutReal atol, rtol, dtFac, max_timestep;
  int mxstep, maximum_order, current_order = 1;
  bool adaptive, adaptive_order, followHighOrder;
  std::deque<State> history;
  std::deque<BoutReal> times;
  ErrorNorm errnorm;

  static BoutReal timestepLimit(BoutReal error, BoutReal tolerance, int order) { // :221-224
    return std::exp(-std::log(error / tolerance) / order);
  }

  /// sum_j coef_j history_j (+ base) for the advance from times[0] to `next` (:189-218)
  State integrate(BoutReal next, int order, const State* base) {
    std::vector<double> c(order);
    for (int j = 0; j < order; ++j) {
      c[j] = abCoefficient(next, times, j, order);
    }
    State update(nv());
    for (size_t i = 0; i < nv(); ++i) {
      std::vector<Term> t;
      if (base != nullptr) {
        t.push_back({1.0, &(*base)[i]});
      }
      for (int j = 0; j < order; ++j) {
        t.push_back({c[j], &history[j][i]});
      }
      update[i] = emptyFrom(*f3d[i].var);
      combine(update[i], t);
    }
    return update;
  }

  // :556-641.  Forms the next state (as handles in `result`) and returns max |full - two-half-steps| of the updates
  BoutReal abStep(BoutReal timeIn, BoutReal dt, int order, const State& cur, State& result) {
    const State full_update = integrate(timeIn + dt, order, nullptr);
    if (!(adaptive && followHighOrder)) {
      form({{1.0, &cur}, {1.0, &full_update}});
      result = current();
    }
    if (!adaptive) {
      return 0.0;
    }
    constexpr BoutReal firstPart = 0.5;
    State half_update = integrate(timeIn + dt * firstPart, order, nullptr);
    times.push_front(timeIn + dt * firstPart);
    if (order == current_order) { // derivative at the half point (:604-621)
      form({{1.0, &cur}, {1.0, &half_update}});
      run_rhs(timeIn + firstPart * dt);
      load(cur);
    } // else: the lower-order trial re-uses the half-point derivative still held by the ddt fields (:598-603)
    history.push_front(derivs());
    half_update = integrate(timeIn + dt, order, &half_update);
    history.pop_front();
    times.pop_front();
    if (followHighOrder) {
      form({{1.0, &cur}, {1.0, &half_update}});
      result = current();
    }
    load(cur);
    return errnorm.maxAbsDiff(full_update, half_update);
  }
};
} // namespace

std::unique_ptr<Solver> Solver::create(Options* opts) {
  Options& o = opts != nullptr ? *opts : Options::root()["solver"];
  const std::string type = o["type"].withDefault(std::string("rk4"));
  if (type == "rk4") {
    if (o["adaptive"].withDefault(false)) {
      return std::make_unique<RK4Adaptive>(&o);
    }
    return std::make_unique<RK4Solver>(&o);
  }
  if (type == "rk3ssp") {
    return std::make_unique<RK3SSP>(&o);
  }
  if (type == "euler") {
    return std::make_unique<EulerSolver>(&o);
  }
  if (type == "rkgeneric") {
    return std::make_unique<RKGenericSolver>(&o);
  }
  if (type == "splitrk") {
    return std::make_unique<SplitRK>(&o);
  }
  if (type == "adams-bashforth") {
    return std::make_unique<AdamsBashforthSolver>(&o);
  }
  throw BoutException("Solver type '{}' is not available: the B200 path implements rk4, rk3ssp, euler, rkgeneric, splitrk and adams-bashforth",
                      type);
}
