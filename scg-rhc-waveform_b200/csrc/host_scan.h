// Host side of the record ingest: the two small text files beside every signal file — the WFDB header `<name>.hea` and
// the JSON side-car `<name>.json` — read and parsed for a whole chunk of records in one call, on a few threads.
//
// The reference reads them one record at a time in Python (wfdb.rdrecord, recordutil.py:137; json.load of the side-car,
// recordutil.py:97-98) — ~55 us of interpreter time per record, more than the record's 2.4 MB take to cross PCIe.  This
// parser covers the COMMON SHAPE of both files and says "not mine" (status 1) for anything else, which the caller then
// sends through the general Python parsers (scgrhc/wfdbio.py, json); it never guesses:
//   header    no comment lines; `<name> <nsig> <fs>[/..] <nsamp>`; every signal line
//             `<file> 16 <gain>(<baseline>)[/<units>] <res> <zero> <init> <checksum> <blocksize> <description>`, one file
//   side-car  a JSON object whose "MacStTime" / "MacEndTime" are "<date> H:M:S" strings and whose "ChamEvents_in_s" is
//             an object of plain numbers with escape-free keys (anything else in the file is skipped structurally)
// What comes out is exactly what the planner consumes (recordutil.py:100-108): the record length in seconds with the date
// ignored, the event times in file order, and the chamber prefix key.split('_')[0] of every event.
#pragma once
#include <atomic>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <sys/stat.h>

#include "scgrhc.h"

namespace scgrhc {
namespace hostscan {

constexpr int kPrefixBytes = 16;

inline bool read_file(const std::string& path, std::string& out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  out.clear();
  char buf[8192];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) {
    out.append(buf, n);
    if (out.size() > (1u << 24)) break;   // not a header / side-car
  }
  const bool ok = !std::ferror(f) && out.size() <= (1u << 24);
  std::fclose(f);
  return ok;
}

// Correctly rounded and independent of the process locale (strtod follows LC_NUMERIC; Python's float() does not).  Values
// out of double range (float() gives inf) and anything from_chars does not take whole are left to the Python parsers.
inline bool to_double(const char* s, const char* t, double& v) {
  const auto r = std::from_chars(s, t, v);
  return r.ec == std::errc() && r.ptr == t;
}

// ---- JSON: a structural skipper plus the three fields ------------------------------------------------
struct Json {
  const char* p;
  const char* e;
  bool bad = false;
  void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  // string at p (opening quote) -> [s, t) without the quotes; `plain` false if it holds an escape or a non-ASCII byte
  bool str(const char*& s, const char*& t, bool& plain) {
    if (p >= e || *p != '"') return false;
    s = ++p;
    plain = true;
    while (p < e && *p != '"') {
      if (*p == '\\') { plain = false; ++p; if (p >= e) return false; }
      else if ((unsigned char)*p < 0x20) return false;
      else if ((unsigned char)*p >= 0x80) plain = false;
      ++p;
    }
    if (p >= e) return false;
    t = p++;
    return true;
  }
  bool number_end(const char*& t) {   // JSON number grammar; t = one past
    const char* q = p;
    if (q < e && *q == '-') ++q;
    if (q >= e) return false;
    if (*q == '0') ++q;
    else if (*q >= '1' && *q <= '9') { while (q < e && *q >= '0' && *q <= '9') ++q; }
    else return false;
    if (q < e && *q == '.') { ++q; if (q >= e || *q < '0' || *q > '9') return false; while (q < e && *q >= '0' && *q <= '9') ++q; }
    if (q < e && (*q == 'e' || *q == 'E')) {
      ++q;
      if (q < e && (*q == '+' || *q == '-')) ++q;
      if (q >= e || *q < '0' || *q > '9') return false;
      while (q < e && *q >= '0' && *q <= '9') ++q;
    }
    t = q;
    return true;
  }
  bool skip(int depth = 0) {          // any value
    ws();
    if (p >= e || depth > 64) return false;
    const char c = *p;
    if (c == '"') { const char *s, *t; bool pl; return str(s, t, pl); }
    if (c == '{' || c == '[') {
      const char close = c == '{' ? '}' : ']';
      ++p; ws();
      if (p < e && *p == close) { ++p; return true; }
      for (;;) {
        if (c == '{') {
          ws();
          const char *s, *t; bool pl;
          if (!str(s, t, pl)) return false;
          ws();
          if (p >= e || *p != ':') return false;
          ++p;
        }
        if (!skip(depth + 1)) return false;
        ws();
        if (p < e && *p == ',') { ++p; continue; }
        if (p < e && *p == close) { ++p; return true; }
        return false;
      }
    }
    if (c == 't') { if (e - p >= 4 && !std::memcmp(p, "true", 4)) { p += 4; return true; } return false; }
    if (c == 'f') { if (e - p >= 5 && !std::memcmp(p, "false", 5)) { p += 5; return true; } return false; }
    if (c == 'n') { if (e - p >= 4 && !std::memcmp(p, "null", 4)) { p += 4; return true; } return false; }
    const char* t;
    if (!number_end(t)) return false;   // NaN / Infinity (json.loads accepts them): not mine
    p = t;
    return true;
  }
};

// "<date> H:M:S" -> seconds of the day, the way `datetime.strptime(s.split()[1], '%H:%M:%S')` reads it
inline bool clock_seconds(const char* s, const char* t, double& out) {
  auto is_sp = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; };
  const char* q = s;
  while (q < t && is_sp(*q)) ++q;
  while (q < t && !is_sp(*q)) ++q;          // first token (the date)
  while (q < t && is_sp(*q)) ++q;
  const char* a = q;
  while (q < t && !is_sp(*q)) ++q;          // second token
  const char* b = q;
  if (a == b) return false;
  int v[3];
  const char* r = a;
  for (int k = 0; k < 3; ++k) {
    int n = 0, d = 0;
    while (r < b && *r >= '0' && *r <= '9' && d < 2) { n = n * 10 + (*r - '0'); ++r; ++d; }
    if (d == 0) return false;
    v[k] = n;
    if (k < 2) { if (r >= b || *r != ':') return false; ++r; }
  }
  if (r != b || v[0] > 23 || v[1] > 59 || v[2] > 59) return false;   // 60 / 61 seconds: strptime's business
  out = (double)(v[0] * 3600 + v[1] * 60 + v[2]);
  return true;
}

struct SideCar {
  bool have_start = false, have_end = false, have_events = false, events_is_object = false;
  double t0 = 0, t1 = 0;
  int n_events = 0;
};

// 0: parsed; 1: not the common shape
inline int parse_sidecar(const std::string& text, int max_events, SideCar& sc, double* ev_time, char* ev_prefix) {
  Json j{text.data(), text.data() + text.size()};
  if (text.size() >= 3 && !std::memcmp(text.data(), "\xef\xbb\xbf", 3)) return 1;
  j.ws();
  if (j.p >= j.e || *j.p != '{') return 1;
  ++j.p; j.ws();
  if (j.p < j.e && *j.p == '}') { ++j.p; }
  else {
    for (;;) {
      j.ws();
      const char *ks, *kt; bool plain;
      if (!j.str(ks, kt, plain)) return 1;
      j.ws();
      if (j.p >= j.e || *j.p != ':') return 1;
      ++j.p; j.ws();
      const size_t kl = (size_t)(kt - ks);
      const bool k_start = plain && kl == 9 && !std::memcmp(ks, "MacStTime", 9);
      const bool k_end = plain && kl == 10 && !std::memcmp(ks, "MacEndTime", 10);
      const bool k_ev = plain && kl == 15 && !std::memcmp(ks, "ChamEvents_in_s", 15);
      if (!plain) return 1;                         // an escaped key could spell one of ours
      if (k_start || k_end) {
        const char *s, *t; bool pl;
        if (j.p >= j.e || *j.p != '"' || !j.str(s, t, pl) || !pl) return 1;
        double sec;
        if (!clock_seconds(s, t, sec)) return 1;
        if (k_start) { sc.t0 = sec; sc.have_start = true; } else { sc.t1 = sec; sc.have_end = true; }
      } else if (k_ev) {
        sc.have_events = true;
        sc.n_events = 0;
        std::vector<std::pair<const char*, size_t>> seen;
        if (j.p < j.e && *j.p == '{') {
          sc.events_is_object = true;
          ++j.p; j.ws();
          if (j.p < j.e && *j.p == '}') { ++j.p; }
          else {
            for (;;) {
              j.ws();
              const char *s, *t; bool pl;
              if (!j.str(s, t, pl) || !pl) return 1;
              j.ws();
              if (j.p >= j.e || *j.p != ':') return 1;
              ++j.p; j.ws();
              const char* ne;
              if (!j.number_end(ne)) return 1;      // strings, null, nested values: float(...) semantics are Python's
              if (sc.n_events >= max_events) return 1;
              const size_t kl2 = (size_t)(t - s);
              if (kl2 == 3 && !std::memcmp(s, "END", 3)) return 1;   // collides with the appended end marker
              size_t pl2 = 0;
              while (pl2 < kl2 && s[pl2] != '_') ++pl2;
              if (pl2 >= (size_t)kPrefixBytes) return 1;
              for (const auto& kv : seen)                             // a repeated key keeps its first position and takes the
                if (kv.second == kl2 && !std::memcmp(kv.first, s, kl2)) return 1;   // last value in a dict: rare, Python's
              seen.emplace_back(s, kl2);
              char* dst = ev_prefix + (size_t)sc.n_events * kPrefixBytes;
              std::memset(dst, 0, kPrefixBytes);
              std::memcpy(dst, s, pl2);
              if (!to_double(j.p, ne, ev_time[sc.n_events])) return 1;
              j.p = ne;
              ++sc.n_events;
              j.ws();
              if (j.p < j.e && *j.p == ',') { ++j.p; continue; }
              if (j.p < j.e && *j.p == '}') { ++j.p; break; }
              return 1;
            }
          }
        } else {
          sc.events_is_object = false;
          if (!j.skip()) return 1;
        }
      } else {
        if (!j.skip()) return 1;
      }
      j.ws();
      if (j.p < j.e && *j.p == ',') { ++j.p; continue; }
      if (j.p < j.e && *j.p == '}') { ++j.p; break; }
      return 1;
    }
  }
  j.ws();
  if (j.p != j.e) return 1;
  if (!sc.have_start || !sc.have_end || !sc.have_events) return 1;   // KeyError territory: Python raises it
  return 0;
}

// ---- WFDB header, common shape -------------------------------------------------------------------------
inline bool parse_int64(const char* s, const char* t, long long& v) {
  if (s == t) return false;
  const char* q = s;
  bool neg = false;
  if (*q == '-' || *q == '+') { neg = *q == '-'; ++q; }
  if (q == t || t - q > 18) return false;
  long long n = 0;
  for (; q < t; ++q) {
    if (*q < '0' || *q > '9') return false;
    n = n * 10 + (*q - '0');
  }
  v = neg ? -n : n;
  return true;
}

// plain decimal / exponent float, nothing Python's float() and from_chars could read differently
inline bool parse_float(const char* s, const char* t, double& v) {
  if (s == t || t - s > 64) return false;
  for (const char* q = s; q < t; ++q)
    if (!((*q >= '0' && *q <= '9') || *q == '.' || *q == '-' || *q == '+' || *q == 'e' || *q == 'E')) return false;
  return to_double(s, t, v);
}

struct Tokens {
  const char* p;
  const char* e;
  bool next(const char*& s, const char*& t) {
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    if (p >= e) return false;
    s = p;
    while (p < e && *p != ' ' && *p != '\t') ++p;
    t = p;
    return true;
  }
};

// 0 ok, 1 not the common shape, 2 another signal count, 3 signal file missing.  expect: nsig_expect descriptions, NUL separated (names_match reports equality)
inline int parse_header(const std::string& text, const std::string& dir, const std::string& rec, int nsig_expect, const char* expect,
                        scgrhc_record_scan& out, double* gains, int32_t* baselines) {
  for (unsigned char c : text)
    if (c >= 0x80 || c == '\r' || c == '\v' || c == '\f') return 1;          // str.split() has more separators than this parser
  size_t pos = 0;
  auto line = [&](const char*& s, const char*& t) {
    if (pos >= text.size()) return false;
    const size_t nl = text.find('\n', pos);
    s = text.data() + pos;
    t = text.data() + (nl == std::string::npos ? text.size() : nl);
    pos = nl == std::string::npos ? text.size() : nl + 1;
    return true;
  };
  const char *ls, *lt;
  if (!line(ls, lt)) return 1;
  if (std::memchr(ls, '#', (size_t)(lt - ls))) return 1;
  Tokens tk{ls, lt};
  const char *s, *t;
  long long nsig = 0, nsamp = 0;
  if (!tk.next(s, t)) return 1;                                     // record name
  if (!tk.next(s, t) || !parse_int64(s, t, nsig) || nsig < 1 || nsig > 64) return 1;
  if (!tk.next(s, t)) return 1;
  {
    const char* sl = (const char*)std::memchr(s, '/', (size_t)(t - s));
    if (!parse_float(s, sl ? sl : t, out.fs)) return 1;
  }
  if (!tk.next(s, t) || !parse_int64(s, t, nsamp) || nsamp < 0) return 1;
  out.nsig = (int32_t)nsig;
  out.names_match = (nsig == nsig_expect) ? 1 : 0;
  if (nsig != nsig_expect) return 2;                                // a different layout: the caller's heterogeneity check
  std::string fname;
  const char* ex = expect;
  for (long long k = 0; k < nsig; ++k) {
    if (!line(ls, lt)) return 1;
    if (ls == lt || *ls == '#') return 1;
    Tokens sk{ls, lt};
    const char *f0, *f1, *g0, *g1;
    if (!sk.next(f0, f1)) return 1;
    if (f0 != ls) return 1;                                         // leading blanks: str.split is fine with them, keep to the plain shape
    if (!sk.next(s, t) || t - s != 2 || s[0] != '1' || s[1] != '6') return 1;
    if (!sk.next(g0, g1)) return 1;
    const char* par = (const char*)std::memchr(g0, '(', (size_t)(g1 - g0));
    if (!par) return 1;
    const char* close = (const char*)std::memchr(par, ')', (size_t)(g1 - par));
    if (!close) return 1;
    double gain;
    long long base;
    if (!parse_float(g0, par, gain) || gain == 0.0 || !parse_int64(par + 1, close, base) || base < INT32_MIN || base > INT32_MAX) return 1;
    for (int skipn = 0; skipn < 5; ++skipn)
      if (!sk.next(s, t)) return 1;                                 // res zero init checksum blocksize
    // description: the rest of the line, stripped (str.split(None, 8)[8].strip())
    const char* d0 = sk.p;
    while (d0 < lt && (*d0 == ' ' || *d0 == '\t')) ++d0;
    const char* d1 = lt;
    while (d1 > d0 && (d1[-1] == ' ' || d1[-1] == '\t')) --d1;
    if (d0 == d1) return 1;
    if (k == 0) fname.assign(f0, f1);
    else if (fname.size() != (size_t)(f1 - f0) || std::memcmp(fname.data(), f0, fname.size())) return 1;
    const size_t el = std::strlen(ex);
    if (el != (size_t)(d1 - d0) || std::memcmp(ex, d0, el)) out.names_match = 0;
    ex += el + 1;
    gains[k] = gain;
    baselines[k] = (int32_t)base;
  }
  if (fname != rec + ".dat") return 1;                               // another signal file name: the general reader resolves it
  struct stat st;
  if (::stat((dir + "/" + fname).c_str(), &st) != 0) return 3;
  const long long on_disk = (long long)st.st_size / (2 * nsig);
  out.rows = nsamp < on_disk ? nsamp : on_disk;
  return 0;
}

inline void scan_one(const std::string& dir, const char* name, int nsig_expect, const char* expect, int max_events,
                     scgrhc_record_scan& out, double* gains, int32_t* baselines, double* ev_time, char* ev_prefix) {
  std::memset(&out, 0, sizeof out);
  out.status = 1;
  const std::string rec(name);
  std::string text;
  if (!read_file(dir + "/" + rec + ".hea", text)) { out.status = 2; return; }
  const int h = parse_header(text, dir, rec, nsig_expect, expect, out, gains, baselines);
  if (h == 3) { out.status = 2; return; }
  if (h == 1) { out.status = 1; return; }
  if (h == 2) { out.status = 0; out.names_match = 0; return; }
  if (!read_file(dir + "/" + rec + ".json", text)) { out.status = 2; return; }
  SideCar sc;
  if (parse_sidecar(text, max_events, sc, ev_time, ev_prefix) != 0) { out.status = 1; return; }
  out.duration_s = sc.t1 - sc.t0;
  out.n_events = sc.events_is_object ? sc.n_events : -1;
  out.status = 0;
}

inline int scan_records(const char* dir, const char* names_blob, int64_t n, const char* expect_blob, int32_t nsig_expect,
                        int32_t max_events, int32_t threads, scgrhc_record_scan* out, double* gains, int32_t* baselines,
                        double* ev_time, char* ev_prefix) {
  if (!dir || n < 0 || (n && (!names_blob || !out || !gains || !baselines || !ev_time || !ev_prefix)) || nsig_expect < 1 ||
      nsig_expect > 64 || max_events < 1 || !expect_blob)
    return SCGRHC_ERR_BAD_ARG;
  std::vector<const char*> names((size_t)n);
  const char* q = names_blob;
  for (int64_t r = 0; r < n; ++r) { names[(size_t)r] = q; q += std::strlen(q) + 1; }
  const std::string d(dir);
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    for (;;) {
      const int64_t r0 = next.fetch_add(16);
      if (r0 >= n) break;
      const int64_t r1 = r0 + 16 < n ? r0 + 16 : n;
      for (int64_t r = r0; r < r1; ++r)
        scan_one(d, names[(size_t)r], nsig_expect, expect_blob, max_events, out[r], gains + r * nsig_expect, baselines + r * nsig_expect,
                 ev_time + r * max_events, ev_prefix + (size_t)r * max_events * kPrefixBytes);
    }
  };
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt > 8) nt = 8;
  if ((int64_t)nt > (n + 15) / 16) nt = (int)((n + 15) / 16);
  if (nt <= 1) { work(); return SCGRHC_OK; }
  std::vector<std::thread> pool;
  try {
    for (int k = 0; k < nt - 1; ++k) pool.emplace_back(work);
  } catch (...) {}
  work();
  for (auto& th : pool) th.join();
  return SCGRHC_OK;
}

}  // namespace hostscan
}  // namespace scgrhc
