// nm_format.cpp -- native '%.4E' text records, byte-identical to the reference's Python '%' formatting
// (write_thrm / write_traj, lammps_remcmc.py:235-256). Host code; C's printf and Python's '%.4E' both
// emit the correctly rounded 5-significant-digit decimal with a two-digit (at least) exponent.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "nm_b200.h"

extern int nm_fail_msg(int code, const char* fmt, ...);

static inline int put_e4(char* dst, double v) {      // " %.4E"
  dst[0] = ' ';
  return 1 + snprintf(dst + 1, 32, "%.4E", v);
}

extern "C" int64_t nm_format_thrm(const double* vals17, char* buf, int64_t cap) {
  if (!vals17) return nm_fail_msg(NM_EINVAL, "nm_format_thrm: null argument");
  char tmp[17 * 40 + 2];
  int n = 0;
  for (int k = 0; k < 17; k++) n += put_e4(tmp + n, vals17[k]);
  tmp[n++] = '\n';
  if (buf) { if (cap < n) return nm_fail_msg(NM_EINVAL, "nm_format_thrm: buffer too small"); memcpy(buf, tmp, n); }
  return n;
}

// '%d %.4E\n' then natoms lines of 3*' %.4E'+'\n'
extern "C" int64_t nm_format_traj(int32_t natoms, double box, const double* x, char* buf, int64_t cap) {
  if (!x || natoms < 0) return nm_fail_msg(NM_EINVAL, "nm_format_traj: bad argument");
  char tmp[160];
  int64_t n = 0;
  int h = snprintf(tmp, sizeof tmp, "%d %.4E\n", natoms, box);
  if (buf) { if (cap < n + h) return nm_fail_msg(NM_EINVAL, "nm_format_traj: buffer too small"); memcpy(buf + n, tmp, h); }
  n += h;
  for (int i = 0; i < natoms; i++) {
    int m = 0;
    for (int a = 0; a < 3; a++) m += put_e4(tmp + m, x[3 * i + a]);
    tmp[m++] = '\n';
    if (buf) { if (cap < n + m) return nm_fail_msg(NM_EINVAL, "nm_format_traj: buffer too small"); memcpy(buf + n, tmp, m); }
    n += m;
  }
  return n;
}

// batch form used by the host driver: nrep trajectories formatted by a pool of threads.
// out_off[nrep+1] receives byte offsets into buf (records are packed back to back in replica order).
extern "C" int64_t nm_format_traj_batch(int32_t nrep, int32_t natoms, const double* box, const double* x,
                                        char* buf, int64_t cap, int64_t* out_off, int32_t nthreads) {
  if (!box || !x || !out_off || nrep < 0) return nm_fail_msg(NM_EINVAL, "nm_format_traj_batch: bad argument");
  const int64_t per_max = 32 + (int64_t)natoms * (3 * 13 + 1) + 64;   // ' -1.2345E+308' is 13 bytes
  std::vector<std::vector<char>> parts(nrep);
  std::vector<int64_t> len(nrep, 0);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nrep) nthreads = nrep > 0 ? nrep : 1;
  auto work = [&](int t) {
    for (int k = t; k < nrep; k += nthreads) {
      parts[k].resize(per_max);
      len[k] = nm_format_traj(natoms, box[k], x + 3 * (size_t)natoms * k, parts[k].data(), per_max);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  int64_t n = 0;
  for (int k = 0; k < nrep; k++) {
    if (len[k] < 0) return len[k];
    out_off[k] = n;
    if (buf) { if (cap < n + len[k]) return nm_fail_msg(NM_EINVAL, "nm_format_traj_batch: buffer too small"); memcpy(buf + n, parts[k].data(), len[k]); }
    n += len[k];
  }
  out_off[nrep] = n;
  return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// N2: streaming output. One call per recorded cycle appends every local replica's record to ITS OWN text file (what
// write_thrm / write_traj do with open(.., 'a'), lammps_remcmc.py:235-256), formatted ONCE by a pool of host threads, so
// the host never holds more than one cycle of text. Optionally the same pass returns what the unmodified parser would
// read back from that text (lammps_parse.py:45,88-93: decimal text -> double -> float32), which is what direct .npy
// emission stores: the '%.4E' round trip keeps 5 significant digits, so float32(x) itself would NOT match.
static inline float parse_back(const char* s) { return (float)strtod(s, nullptr); }

static int64_t append_file(const char* path, const char* data, size_t n) {
  FILE* f = fopen(path, "ab");
  if (!f) return nm_fail_msg(NM_EINVAL, "cannot open %s for appending", path);
  const size_t w = fwrite(data, 1, n, f);
  if (fclose(f) != 0 || w != n) return nm_fail_msg(NM_EINVAL, "short write to %s", path);
  return (int64_t)n;
}

extern "C" int64_t nm_append_traj_batch(int32_t nrep, int32_t natoms, const double* box, const double* x,
                                        const char* const* paths, int32_t nthreads, float* pos_out, float* box_out) {
  if (!box || !x || !paths || nrep < 0 || natoms < 0) return nm_fail_msg(NM_EINVAL, "nm_append_traj_batch: bad argument");
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nrep) nthreads = nrep > 0 ? nrep : 1;
  std::vector<int64_t> len(nrep, 0);
  auto work = [&](int t) {
    std::vector<char> buf(64 + (size_t)natoms * (3 * 24 + 1));
    for (int k = t; k < nrep; k += nthreads) {
      char* p = buf.data();
      p += snprintf(p, 64, "%d %.4E\n", natoms, box[k]);
      if (box_out) { char tmp[40]; snprintf(tmp, sizeof tmp, "%.4E", box[k]); box_out[k] = parse_back(tmp); }
      const double* xk = x + 3 * (size_t)natoms * k;
      for (int i = 0; i < 3 * natoms; i++) {
        *p++ = ' ';
        const int m = snprintf(p, 24, "%.4E", xk[i]);
        if (pos_out) pos_out[3 * (size_t)natoms * k + i] = parse_back(p);
        p += m;
        if (i % 3 == 2) *p++ = '\n';
      }
      len[k] = paths[k] ? append_file(paths[k], buf.data(), (size_t)(p - buf.data())) : (int64_t)(p - buf.data());
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  int64_t n = 0;
  for (int k = 0; k < nrep; k++) { if (len[k] < 0) return len[k]; n += len[k]; }
  return n;
}

// vals: [nrep][17] in the column order of write_thrm; parsed_out: optional float32 [nrep][17] (np.loadtxt(dtype=float32))
extern "C" int64_t nm_append_thrm_batch(int32_t nrep, const double* vals, const char* const* paths, float* parsed_out) {
  if (!vals || !paths || nrep < 0) return nm_fail_msg(NM_EINVAL, "nm_append_thrm_batch: bad argument");
  int64_t n = 0;
  for (int k = 0; k < nrep; k++) {
    char tmp[17 * 40 + 2];
    int m = 0;
    for (int c = 0; c < 17; c++) {
      tmp[m++] = ' ';
      const int w = snprintf(tmp + m, 32, "%.4E", vals[17 * (size_t)k + c]);
      if (parsed_out) parsed_out[17 * (size_t)k + c] = parse_back(tmp + m);
      m += w;
    }
    tmp[m++] = '\n';
    if (paths[k]) { const int64_t r = append_file(paths[k], tmp, (size_t)m); if (r < 0) return r; }
    n += m;
  }
  return n;
}
