// Output step of the device-resident solver: what PhysicsModelMonitor::call does in the reference
// (src/physics/physicsmodel.cxx:241-268 -> Solver::outputVars, src/solver/solver.cxx:694-724 -> OptionsNetCDF),
// re-designed for fields that live in HBM:
//   * a snapshot is a set of copy-on-write handles (no device copy; later steps write to fresh blocks),
//   * it leaves the device on a second stream into page-locked staging while the next output interval computes,
//   * a host thread converts to big-endian and appends the record to the dump file, one interval later still.
// The file is NetCDF classic with 64-bit offsets (CDF-2): dimensions t (record), x, y, z; the integer scalars
// boutdata.collect needs (NXPE, NYPE, MXSUB, MYSUB, MXG, MYG, MZ, nx, ny, nz); t_array(t), iteration(t) and one
// double variable (t, x, y, z) per evolving or model-registered field -- the layout of BOUT.dmp.<rank>.nc.
#include "bout/b200_host.hxx"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>

namespace bout {
namespace b200 {
namespace {
void put32(std::vector<unsigned char>& b, std::int32_t v) {
  const auto u = static_cast<std::uint32_t>(v);
  for (int s = 24; s >= 0; s -= 8) {
    b.push_back(static_cast<unsigned char>((u >> s) & 0xffu));
  }
}
void put64(std::vector<unsigned char>& b, std::int64_t v) {
  const auto u = static_cast<std::uint64_t>(v);
  for (int s = 56; s >= 0; s -= 8) {
    b.push_back(static_cast<unsigned char>((u >> s) & 0xffu));
  }
}
void putName(std::vector<unsigned char>& b, const std::string& n) {
  put32(b, static_cast<std::int32_t>(n.size()));
  b.insert(b.end(), n.begin(), n.end());
  while (b.size() % 4 != 0) {
    b.push_back(0);
  }
}
constexpr std::int32_t NC_DIMENSION = 10, NC_VARIABLE = 11, NC_INT = 4, NC_DOUBLE = 6;

struct VarDesc {
  std::string name;
  std::vector<int> dims; // dimension ids
  std::int32_t type;
  std::int64_t vsize;    // bytes (one record for record variables)
  std::int64_t begin = 0;
};

std::uint64_t swap64(std::uint64_t v) { return __builtin_bswap64(v); }
} // namespace

struct DumpFile::Impl {
  std::FILE* fp = nullptr;
  std::string path;
  std::int64_t rec_begin = 0, rec_size = 0;
  std::int32_t numrecs = 0;
  size_t field_count = 0;
  std::vector<std::int64_t> field_offset; // within a record
  std::int64_t t_offset = 0, iter_offset = 0;
  std::vector<unsigned char> scratch;
};

DumpFile::DumpFile(const std::string& path, int nx, int ny, int nz, const std::vector<DumpScalar>& scalars,
                   const std::vector<DumpOnceField>& once, const std::vector<std::string>& fields)
    : impl(new Impl) {
  impl->path = path;
  impl->field_count = static_cast<size_t>(nx) * ny * nz;
  std::vector<VarDesc> vars;
  for (const auto& s : scalars) {
    vars.push_back({s.name, {}, s.is_double ? NC_DOUBLE : NC_INT, s.is_double ? 8 : 4});
  }
  for (const auto& f : once) { // written once (SAVE_ONCE of a Field2D / Field3D): (x, y) or (x, y, z), no record dimension
    const size_t n = static_cast<size_t>(nx) * ny * (f.is3d ? nz : 1);
    if (f.data.size() != n) {
      throw BoutException("DumpFile: '{}' has {} values, expected {}", f.name, f.data.size(), n);
    }
    vars.push_back({f.name, f.is3d ? std::vector<int>{1, 2, 3} : std::vector<int>{1, 2}, NC_DOUBLE,
                    static_cast<std::int64_t>(n) * 8});
  }
  const size_t first_record_var = vars.size();
  vars.push_back({"t_array", {0}, NC_DOUBLE, 8});
  vars.push_back({"iteration", {0}, NC_INT, 4});
  for (const auto& f : fields) {
    vars.push_back({f, {0, 1, 2, 3}, NC_DOUBLE, static_cast<std::int64_t>(impl->field_count) * 8});
  }
  // header size is needed for the offsets, and the offsets are in the header: lay it out twice
  std::vector<unsigned char> h;
  for (int pass = 0; pass < 2; ++pass) {
    const std::int64_t header_size = static_cast<std::int64_t>(h.size());
    if (pass == 1) {
      std::int64_t off = header_size;
      for (size_t i = 0; i < first_record_var; ++i) {
        vars[i].begin = off;
        off += vars[i].vsize;
      }
      impl->rec_begin = off;
      std::int64_t in_rec = 0;
      for (size_t i = first_record_var; i < vars.size(); ++i) {
        vars[i].begin = off + in_rec;
        if (vars[i].name == "t_array") {
          impl->t_offset = in_rec;
        } else if (vars[i].name == "iteration") {
          impl->iter_offset = in_rec;
        } else {
          impl->field_offset.push_back(in_rec);
        }
        in_rec += vars[i].vsize;
      }
      impl->rec_size = in_rec;
    }
    h.clear();
    h.insert(h.end(), {'C', 'D', 'F', 2});
    put32(h, 0); // numrecs, rewritten after every record
    put32(h, NC_DIMENSION);
    put32(h, 4);
    const std::pair<const char*, int> dims[4] = {{"t", 0}, {"x", nx}, {"y", ny}, {"z", nz}};
    for (const auto& d : dims) {
      putName(h, d.first);
      put32(h, d.second);
    }
    put32(h, 0); // no global attributes
    put32(h, 0);
    put32(h, NC_VARIABLE);
    put32(h, static_cast<std::int32_t>(vars.size()));
    for (const auto& v : vars) {
      putName(h, v.name);
      put32(h, static_cast<std::int32_t>(v.dims.size()));
      for (int d : v.dims) {
        put32(h, d);
      }
      put32(h, 0); // no attributes
      put32(h, 0);
      put32(h, v.type);
      put32(h, static_cast<std::int32_t>(v.vsize > 0x7fffffff ? 0x7fffffff : v.vsize)); // vsize is advisory beyond 2 GiB
      put64(h, v.begin);
    }
  }
  impl->fp = std::fopen(path.c_str(), "wb+");
  if (impl->fp == nullptr) {
    throw BoutException("Could not open output file '{}'", path);
  }
  auto putDouble = [&](double v) {
    std::uint64_t u;
    std::memcpy(&u, &v, 8);
    put64(h, static_cast<std::int64_t>(u));
  };
  for (const auto& s : scalars) {
    if (s.is_double) {
      putDouble(s.d);
    } else {
      put32(h, s.i);
    }
  }
  for (const auto& f : once) {
    for (double v : f.data) {
      putDouble(v);
    }
  }
  if (std::fwrite(h.data(), 1, h.size(), impl->fp) != h.size()) {
    throw BoutException("Short write to '{}'", path);
  }
  std::fflush(impl->fp);
}

DumpFile::~DumpFile() {
  if (impl->fp != nullptr) {
    std::fclose(impl->fp);
  }
}

int DumpFile::records() const { return impl->numrecs; }

void DumpFile::append(double t, int iteration, const std::vector<const double*>& host_fields) {
  if (host_fields.size() != impl->field_offset.size()) {
    throw BoutException("DumpFile::append: {} fields given, {} declared", host_fields.size(), impl->field_offset.size());
  }
  const std::int64_t base = impl->rec_begin + static_cast<std::int64_t>(impl->numrecs) * impl->rec_size;
  auto writeAt = [&](std::int64_t off, const void* p, size_t n) {
    if (fseeko(impl->fp, static_cast<off_t>(off), SEEK_SET) != 0 || std::fwrite(p, 1, n, impl->fp) != n) {
      throw BoutException("Write to '{}' failed", impl->path);
    }
  };
  std::uint64_t tb;
  std::memcpy(&tb, &t, 8);
  tb = swap64(tb);
  writeAt(base + impl->t_offset, &tb, 8);
  std::vector<unsigned char> ib;
  put32(ib, iteration);
  writeAt(base + impl->iter_offset, ib.data(), 4);
  impl->scratch.resize(impl->field_count * 8);
  for (size_t f = 0; f < host_fields.size(); ++f) {
    const auto* src = reinterpret_cast<const std::uint64_t*>(host_fields[f]);
    auto* dst = reinterpret_cast<std::uint64_t*>(impl->scratch.data());
    for (size_t i = 0; i < impl->field_count; ++i) {
      dst[i] = swap64(src[i]);
    }
    writeAt(base + impl->field_offset[f], impl->scratch.data(), impl->scratch.size());
  }
  ++impl->numrecs;
  std::vector<unsigned char> nb;
  put32(nb, impl->numrecs);
  writeAt(4, nb.data(), 4);
  std::fflush(impl->fp);
}

// ---- reading the non-record variables of such a file back (restart) ------------------------------------------
namespace {
struct Cursor {
  const std::vector<unsigned char>& b;
  size_t pos = 0;
  std::int64_t get(int bytes) {
    if (pos + bytes > b.size()) {
      throw BoutException("NetCDF header is truncated");
    }
    std::uint64_t v = 0;
    for (int i = 0; i < bytes; ++i) {
      v = (v << 8) | b[pos++];
    }
    return bytes == 4 ? static_cast<std::int64_t>(static_cast<std::int32_t>(v)) : static_cast<std::int64_t>(v);
  }
  std::string name() {
    const auto n = static_cast<size_t>(get(4));
    if (pos + n > b.size()) {
      throw BoutException("NetCDF header is truncated");
    }
    std::string s(b.begin() + pos, b.begin() + pos + n);
    pos += (n + 3) / 4 * 4;
    return s;
  }
  void skipAttributes() { // tag + count, then name, type, nelems, values padded to 4 bytes
    const auto tag = get(4);
    const auto n = get(4);
    if (tag == 0 && n == 0) {
      return;
    }
    static const int size_of[] = {0, 1, 1, 2, 4, 4, 8};
    for (std::int64_t i = 0; i < n; ++i) {
      name();
      const auto type = get(4);
      const auto nel = get(4);
      if (type < 1 || type > 6) {
        throw BoutException("NetCDF attribute of unknown type {}", type);
      }
      pos += (static_cast<size_t>(nel) * size_of[type] + 3) / 4 * 4;
    }
  }
};
} // namespace

std::map<std::string, NcVariable> readNcFixedVariables(const std::string& path) {
  std::FILE* fp = std::fopen(path.c_str(), "rb");
  if (fp == nullptr) {
    throw BoutException("Could not open '{}'", path);
  }
  std::vector<unsigned char> all;
  unsigned char chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), fp)) > 0) {
    all.insert(all.end(), chunk, chunk + got);
  }
  std::fclose(fp);
  if (all.size() < 8 || all[0] != 'C' || all[1] != 'D' || all[2] != 'F' || (all[3] != 1 && all[3] != 2)) {
    throw BoutException("'{}' is not a NetCDF classic file", path);
  }
  const int offset_bytes = all[3] == 2 ? 8 : 4;
  Cursor c{all, 4};
  c.get(4); // numrecs
  std::vector<std::int64_t> dim_len;
  {
    const auto tag = c.get(4);
    const auto n = c.get(4);
    if (tag != 0 || n != 0) {
      for (std::int64_t i = 0; i < n; ++i) {
        c.name();
        dim_len.push_back(c.get(4));
      }
    }
  }
  c.skipAttributes();
  std::map<std::string, NcVariable> out;
  const auto tag = c.get(4);
  const auto nvars = c.get(4);
  if (tag == 0 && nvars == 0) {
    return out;
  }
  for (std::int64_t i = 0; i < nvars; ++i) {
    const std::string name = c.name();
    const auto nd = c.get(4);
    NcVariable v;
    bool record = false;
    size_t count = 1;
    for (std::int64_t d = 0; d < nd; ++d) {
      const auto id = c.get(4);
      const auto len = dim_len.at(static_cast<size_t>(id));
      record = record || len == 0;
      v.shape.push_back(static_cast<int>(len));
      count *= static_cast<size_t>(len);
    }
    c.skipAttributes();
    const auto type = c.get(4);
    c.get(4); // vsize
    const auto begin = static_cast<size_t>(c.get(offset_bytes));
    if (record || (type != NC_INT && type != NC_DOUBLE)) {
      continue; // record variables and other types are not needed for a restart
    }
    const size_t width = type == NC_INT ? 4 : 8;
    if (begin + count * width > all.size()) {
      throw BoutException("'{}': variable '{}' extends past the end of the file", path, name);
    }
    v.is_int = type == NC_INT;
    v.data.resize(count);
    Cursor d{all, begin};
    for (size_t k = 0; k < count; ++k) {
      if (v.is_int) {
        v.data[k] = static_cast<double>(d.get(4));
      } else {
        const auto bits = static_cast<std::uint64_t>(d.get(8));
        std::memcpy(&v.data[k], &bits, 8);
      }
    }
    out[name] = std::move(v);
  }
  return out;
}

// Restart file: the evolving variables at the latest output plus tt (time) and hist_hi (iteration), as
// Solver::outputVars(restart_options, false) (solver.cxx:694-724); written to <path>.tmp and renamed.
void writeRestartFile(const std::string& path, int nx, int ny, int nz, double time, int iteration,
                      const std::vector<std::string>& names, const std::vector<const double*>& host_fields) {
  std::vector<DumpOnceField> once;
  const size_t count = static_cast<size_t>(nx) * ny * nz;
  for (size_t f = 0; f < names.size(); ++f) {
    once.push_back({names[f], true, std::vector<double>(host_fields[f], host_fields[f] + count)});
  }
  const std::string tmp = path + ".tmp";
  {
    DumpFile file(tmp, nx, ny, nz, {{"hist_hi", false, iteration, 0.0}, {"tt", true, 0, time}}, once, {});
  }
  if (std::rename(tmp.c_str(), path.c_str()) != 0) {
    throw BoutException("Could not move '{}' to '{}'", tmp, path);
  }
}

// ---- the pipeline ------------------------------------------------------------------------------------------
OutputPipeline::OutputPipeline(const std::string& path, Mesh* mesh, const std::vector<std::string>& names,
                               std::vector<DumpScalar> extra_scalars, const std::vector<DumpOnceField>& once,
                               const std::string& restart_path_in, size_t n_evolving_in)
    : nfields(names.size()), restart_path(restart_path_in), n_evolving(n_evolving_in), field_names(names) {
  const auto g = mesh->grid();
  dims[0] = g.nx, dims[1] = g.ny, dims[2] = g.nz;
  count = static_cast<size_t>(g.nx) * g.ny * g.nz;
  std::vector<DumpScalar> scalars;
  const std::pair<const char*, int> ints[] = {
      {"NXPE", mesh->NXPE}, {"NYPE", mesh->NYPE}, {"MXSUB", g.xend - g.xstart + 1}, {"MYSUB", g.yend - g.ystart + 1},
      {"MXG", mesh->MXG}, {"MYG", mesh->MYG}, {"MZ", g.nz}, {"nx", mesh->GlobalNx}, {"ny", mesh->GlobalNy - 2 * mesh->MYG},
      {"nz", g.nz}, {"PE_XIND", mesh->PE_XIND}, {"PE_YIND", mesh->PE_YIND}}; // global sizes (boutdata.collect's stitching)
  for (const auto& kv : ints) {
    scalars.push_back({kv.first, false, kv.second, 0.0});
  }
  scalars.insert(scalars.end(), extra_scalars.begin(), extra_scalars.end());
  if (!path.empty()) {
    file = std::make_unique<DumpFile>(path, g.nx, g.ny, g.nz, scalars, once, names);
  }
  check(bout_b200_stream_create(&copy_stream));
  for (auto& s : slots) {
    void* p = nullptr;
    check(bout_b200_host_alloc(&p, nfields * count * sizeof(double)));
    s.staging = static_cast<double*>(p);
  }
}

OutputPipeline::~OutputPipeline() {
  try {
    finish();
  } catch (...) { // destructors do not throw; finish() is called explicitly on the normal path
  }
  for (auto& s : slots) {
    if (s.staging != nullptr) {
      bout_b200_host_free(s.staging);
    }
  }
  if (copy_stream != nullptr) {
    bout_b200_stream_destroy(copy_stream);
  }
}

void OutputPipeline::retire(Slot& s) { // the record's host thread has finished with the slot
  if (s.writing.valid()) {
    s.writing.get(); // rethrows a failed write
  }
}

void OutputPipeline::handOver(Slot& s) { // device -> host copy done: release the device blocks, start the disk write
  if (!s.copying) {
    return;
  }
  check(bout_b200_stream_sync(copy_stream));
  s.held.clear();
  s.copying = false;
  std::vector<const double*> ptrs(nfields);
  for (size_t f = 0; f < nfields; ++f) {
    ptrs[f] = s.staging + f * count;
  }
  DumpFile* out = file.get();
  const double t = s.time;
  const int it = s.iteration;
  s.writing = std::async(std::launch::async, [this, out, t, it, ptrs]() {
    if (out != nullptr) {
      out->append(t, it, ptrs);
    }
    if (!restart_path.empty()) { // the evolving variables come first in every snapshot
      const std::vector<std::string> names(field_names.begin(), field_names.begin() + n_evolving);
      const std::vector<const double*> fields(ptrs.begin(), ptrs.begin() + n_evolving);
      writeRestartFile(restart_path, dims[0], dims[1], dims[2], t, it, names, fields);
    }
  });
}

void OutputPipeline::snapshot(double time, int iteration, const std::vector<const Field3D*>& fields) {
  if (fields.size() != nfields) {
    throw BoutException("OutputPipeline::snapshot: {} fields given, {} declared", fields.size(), nfields);
  }
  Slot& cur = slots[next % 2];
  Slot& prev = slots[(next + 1) % 2];
  retire(cur);    // record k-2 (same slot) is on disk; also keeps the records in order
  handOver(prev); // record k-1 has left the device: write it while interval k computes
  cur.time = time;
  cur.iteration = iteration;
  cur.held.clear();
  // device pointers first: asking for one launches a field's pending (deferred) combination on the compute stream,
  // and that launch has to be ahead of the point the copy stream waits for
  std::vector<const double*> src(nfields);
  for (size_t f = 0; f < nfields; ++f) {
    if (!fields[f]->isAllocated()) {
      throw BoutException("OutputPipeline::snapshot: field {} has no data", f);
    }
    cur.held.push_back(*fields[f]); // copy-on-write handle: later steps cannot overwrite this block
    src[f] = fields[f]->dptr();
  }
  check(bout_b200_stream_wait(copy_stream, nullptr)); // the snapshot is complete once the compute stream gets here
  for (size_t f = 0; f < nfields; ++f) {
    check(bout_b200_memcpy_d2h_async(cur.staging + f * count, src[f], count * sizeof(double), copy_stream));
  }
  cur.copying = true;
  ++next;
}

void OutputPipeline::finish() {
  for (int k = 0; k < 2; ++k) { // oldest first
    Slot& s = slots[(next + k) % 2];
    if (s.copying) {
      retire(slots[(next + k + 1) % 2]); // the other slot's write precedes this record
      handOver(s);
    }
  }
  for (auto& s : slots) {
    retire(s);
  }
}

int OutputPipeline::recordsWritten() const { return file ? file->records() : 0; }

} // namespace b200
} // namespace bout
