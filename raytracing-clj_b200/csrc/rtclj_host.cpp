// rtclj_host.cpp -- host-only entry points of the C ABI: the steps either side of the
// render loop: quantisation + P3 encoding after it, camera derivation
// and scene generation before it.  Plain C++17, IEEE double, compiled with
// -ffp-contract=off so the arithmetic is what the JVM (and camera.py) computes.
#include "../../include/rtclj_b200.h"
#include "rtclj_error.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include <thread>

// ---- the thread-local error message of the whole library (rtclj_last_error)
namespace { thread_local std::string g_err; }
int rtclj_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
void rtclj_error_set(const char* message) { g_err = message ? message : ""; }
const char* rtclj_error_get() { return g_err.c_str(); }
extern "C" const char* rtclj_last_error(void) { return g_err.c_str(); }

namespace {

struct V { double x, y, z; };
inline V mk(double x, double y, double z) { return V{x, y, z}; }
inline V ld(const double* p) { return V{p[0], p[1], p[2]}; }
inline void st(double* p, V v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline V add(V a, V b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V sub(V a, V b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V muls(V a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V divs(V a, double s) { return mk(a.x / s, a.y / s, a.z / s); }
inline V neg(V a) { return mk(-a.x, -a.y, -a.z); }
inline V cross(V u, V v) {  // vec3a.clj:64-67
  return mk(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
inline double length(V a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V unit(V a) { return divs(a, length(a)); }
inline double deg_to_rad(double d) { return d * M_PI / 180.0; }  // raytracing.clj:60-61

int64_t gcd64(int64_t a, int64_t b) {
  while (b) { int64_t t = a % b; a = b; b = t; }
  return a < 0 ? -a : a;
}

inline int quantise(double c, bool linear) {  // write-color!, raytracing.clj:19-26
  double v;
  if (linear) {
    v = 255.999 * c;
  } else {
    double g = c > 0.0 ? std::sqrt(c) : 0.0;
    double lo = g > 0.0 ? g : 0.0;
    double cl = lo < 0.999 ? lo : 0.999;
    v = 256.0 * cl;
  }
  return (v != v) ? 0 : (int)v;
}

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace

extern "C" {

int rtclj_quantise_rgb8(const double* linear, size_t n_values, uint32_t flags, uint8_t* out) {
  if ((!linear || !out) && n_values) return rtclj_fail(RTCLJ_E_INVALID, "rtclj_quantise_rgb8: null buffer");
  const bool lin = (flags & RTCLJ_F_QUANT_LINEAR) != 0;
  for (size_t i = 0; i < n_values; ++i) out[i] = (uint8_t)quantise(linear[i], lin);
  return RTCLJ_OK;
}

int rtclj_encode_ppm_p3(const uint8_t* rgb8, int32_t width, int32_t height, char* out, size_t capacity,
                        size_t* len) {
  if (width <= 0 || height <= 0 || !len) return rtclj_fail(RTCLJ_E_INVALID, "image size must be positive and len non-null (%d x %d)", width, height);
  char header[64];
  const int hl = std::snprintf(header, sizeof header, "P3\n%d %d\n255\n", width, height);
  const size_t npix = (size_t)width * (size_t)height;
  if (!out) {  // sizing call: worst case "255 255 255\n" (+4: room for the one-pass writer's wide copies)
    *len = (size_t)hl + npix * 12 + 4;
    return RTCLJ_OK;
  }
  if (!rgb8) return rtclj_fail(RTCLJ_E_INVALID, "null image");
  // "ddd" + separator slot, and the digit count, per byte value
  static const struct Lut { char s[256][4]; unsigned char n[256]; Lut() {
      for (int v = 0; v < 256; ++v) { n[v] = (unsigned char)std::snprintf(s[v], 4, "%d", v); s[v][n[v]] = ' '; }
    } } lut;
  const size_t worst = (size_t)hl + npix * 12;
  size_t need = worst;
  // Large images (the 8.3 M lines of a 3840x2160 scene.ppm): several threads, each on a stretch of pixels -- one pass
  // to measure the stretches, one to write them at their offsets; the bytes are those of the loop below.
  if (npix >= ((size_t)1 << 18)) {
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned T = std::max(1u, std::min((hw ? hw : 1u) / 2u, 16u));  // half of the cores: the one-thread loop already runs at 7 GB/s
    if (T >= 2) {
      std::vector<size_t> bytes(T, 0), first(T + 1);
      for (unsigned k = 0; k <= T; ++k) first[k] = npix / T * k + std::min<size_t>(k, npix % T);
      auto run = [&](auto&& fn) {
        std::vector<std::thread> th;
        try { for (unsigned k = 1; k < T; ++k) th.emplace_back(fn, k); } catch (...) { for (auto& t : th) t.join(); throw; }
        fn(0u);
        for (auto& t : th) t.join();
      };
      try {
        run([&](unsigned k) {
          size_t b = 0;  // (a private sum: the shared array is written once)
          for (const uint8_t* q = rgb8 + 3 * first[k], * const e = rgb8 + 3 * first[k + 1]; q < e; ++q) b += lut.n[*q] + 1u;
          bytes[k] = b;
        });
        need = (size_t)hl;
        std::vector<size_t> off(T);
        for (unsigned k = 0; k < T; ++k) { off[k] = need; need += bytes[k]; }
        *len = need;
        if (need > capacity) return rtclj_fail(RTCLJ_E_BUFFER, "P3 text needs %zu bytes, capacity is %zu", need, capacity);
        std::memcpy(out, header, (size_t)hl);
        run([&](unsigned k) {
          char* w = out + off[k];
          const uint8_t* src = rgb8 + 3 * first[k];
          const size_t n = first[k + 1] - first[k], safe = n > 2 ? n - 2 : 0;  // the 4-byte copies must not reach into the next stretch
          for (size_t p = 0; p < safe; ++p, src += 3) {
            const unsigned r = src[0], g = src[1], b = src[2];
            std::memcpy(w, lut.s[r], 4); w += lut.n[r] + 1u;
            std::memcpy(w, lut.s[g], 4); w += lut.n[g] + 1u;
            std::memcpy(w, lut.s[b], 4); w += lut.n[b];
            *w++ = '\n';
          }
          for (size_t p = safe; p < n; ++p, src += 3)
            for (int ch = 0; ch < 3; ++ch) {
              const unsigned v = src[ch], m = lut.n[v];
              for (unsigned i = 0; i < m; ++i) *w++ = lut.s[v][i];
              *w++ = ch == 2 ? '\n' : ' ';
            }
        });
        return RTCLJ_OK;
      } catch (...) {
        // no threads to be had: the sequential writer below
      }
    }
  }
  if (capacity < worst + 4) {  // not provably large enough: measure first
    need = (size_t)hl;
    for (size_t i = 0; i < npix * 3; ++i) need += lut.n[rgb8[i]] + 1u;
    *len = need;
    if (need > capacity) return rtclj_fail(RTCLJ_E_BUFFER, "P3 text needs %zu bytes, capacity is %zu", need, capacity);
  }
  std::memcpy(out, header, (size_t)hl);
  char* w = out + hl;
  // 4-byte copies of "ddd " overrun the text by at most 3 bytes; the last pixels go bytewise
  const size_t safe = (capacity >= worst + 4) ? npix : (npix > 2 ? npix - 2 : 0);
  const uint8_t* src = rgb8;
  for (size_t p = 0; p < safe; ++p, src += 3) {
    const unsigned r = src[0], g = src[1], b = src[2];
    std::memcpy(w, lut.s[r], 4); w += lut.n[r] + 1u;
    std::memcpy(w, lut.s[g], 4); w += lut.n[g] + 1u;
    std::memcpy(w, lut.s[b], 4); w += lut.n[b];
    *w++ = '\n';
  }
  for (size_t p = safe; p < npix; ++p, src += 3) {
    for (int ch = 0; ch < 3; ++ch) {
      const unsigned v = src[ch], n = lut.n[v];
      for (unsigned k = 0; k < n; ++k) *w++ = lut.s[v][k];
      *w++ = ch == 2 ? '\n' : ' ';
    }
  }
  *len = (size_t)(w - out);
  return RTCLJ_OK;
}

// ---- PNG (RGB, 8 bit, no interlace) with stored deflate blocks: signature, IHDR, IDAT, IEND.
namespace {
struct Crc32 {  // slicing-by-8 over the reflected polynomial 0xEDB88320
  uint32_t table[8][256];
  Crc32() {
    for (uint32_t n = 0; n < 256; ++n) {
      uint32_t c = n;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; ++n)
      for (int t = 1; t < 8; ++t) table[t][n] = table[0][table[t - 1][n] & 0xffu] ^ (table[t - 1][n] >> 8);
  }
  uint32_t run(uint32_t crc, const uint8_t* p, size_t n) const {
    while (n >= 8) {
      uint32_t lo, hi;
      std::memcpy(&lo, p, 4); std::memcpy(&hi, p + 4, 4);  // little-endian hosts (x86-64, aarch64)
      lo ^= crc;
      crc = table[7][lo & 0xffu] ^ table[6][(lo >> 8) & 0xffu] ^ table[5][(lo >> 16) & 0xffu] ^ table[4][lo >> 24] ^
            table[3][hi & 0xffu] ^ table[2][(hi >> 8) & 0xffu] ^ table[1][(hi >> 16) & 0xffu] ^ table[0][hi >> 24];
      p += 8; n -= 8;
    }
    for (size_t i = 0; i < n; ++i) crc = table[0][(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
  }
};
struct Adler32 {  // sums reduced every 5552 bytes, the longest run that cannot overflow 32 bits
  uint32_t a = 1, b = 0;
  void run(const uint8_t* p, size_t n) {
    while (n) {
      const size_t k = n < 5552 ? n : 5552;
      for (size_t i = 0; i < k; ++i) { a += p[i]; b += a; }
      a %= 65521u; b %= 65521u;
      p += k; n -= k;
    }
  }
};
inline void be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
}  // namespace

int rtclj_encode_png(const uint8_t* rgb8, int32_t width, int32_t height, uint8_t* out, size_t capacity,
                     size_t* len) {
  if (width <= 0 || height <= 0 || !len) return rtclj_fail(RTCLJ_E_INVALID, "image size must be positive and len non-null (%d x %d)", width, height);
  const size_t row = (size_t)width * 3 + 1;          // filter byte + pixels
  const size_t raw = row * (size_t)height;           // bytes handed to deflate
  const size_t nblocks = (raw + 65534) / 65535;      // stored blocks of <= 65535 bytes
  const size_t zlen = 2 + raw + 5 * nblocks + 4;     // zlib header, blocks, adler32
  const size_t need = 8 + (12 + 13) + (12 + zlen) + 12;
  *len = need;
  if (!out) return RTCLJ_OK;
  if (!rgb8) return rtclj_fail(RTCLJ_E_INVALID, "null image");
  if (need > capacity) return rtclj_fail(RTCLJ_E_BUFFER, "PNG needs %zu bytes, capacity is %zu", need, capacity);
  if (zlen > 0xffffffffull) return rtclj_fail(RTCLJ_E_INVALID, "image too large for one PNG IDAT chunk");
  static const Crc32 crc;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  uint8_t* w = out;
  std::memcpy(w, sig, 8); w += 8;
  // IHDR
  be32(w, 13); std::memcpy(w + 4, "IHDR", 4);
  be32(w + 8, (uint32_t)width); be32(w + 12, (uint32_t)height);
  w[16] = 8; w[17] = 2; w[18] = 0; w[19] = 0; w[20] = 0;  // 8 bit, truecolour, deflate, no filter method, no interlace
  be32(w + 21, crc.run(0xffffffffu, w + 4, 17) ^ 0xffffffffu);
  w += 25;
  // IDAT
  be32(w, (uint32_t)zlen); std::memcpy(w + 4, "IDAT", 4);
  uint8_t* z = w + 8;
  z[0] = 0x78; z[1] = 0x01;
  uint8_t* q = z + 2;
  // the raw stream (filter byte 0 + the row, per row) cut into stored blocks of <= 65535 bytes
  Adler32 adler;
  size_t produced = 0, in_block = 0;
  auto emit = [&](const uint8_t* p, size_t n) {
    adler.run(p, n);
    while (n) {
      if (in_block == 0) {
        const size_t left = raw - produced, m = left < 65535 ? left : 65535;
        q[0] = (produced + m == raw) ? 1 : 0;  // BFINAL, BTYPE = 00 (stored)
        q[1] = (uint8_t)(m & 0xff); q[2] = (uint8_t)(m >> 8); q[3] = (uint8_t)~q[1]; q[4] = (uint8_t)~q[2];
        q += 5;
        in_block = m;
      }
      const size_t k = n < in_block ? n : in_block;
      std::memcpy(q, p, k);
      q += k; p += k; n -= k; in_block -= k; produced += k;
    }
  };
  static const uint8_t filter_none = 0;
  for (int32_t j = 0; j < height; ++j) {
    emit(&filter_none, 1);
    emit(rgb8 + (size_t)j * (size_t)width * 3, (size_t)width * 3);
  }
  be32(q, (adler.b << 16) | adler.a); q += 4;
  be32(q, crc.run(0xffffffffu, w + 4, 4 + zlen) ^ 0xffffffffu);
  w = q + 4;
  // IEND
  be32(w, 0); std::memcpy(w + 4, "IEND", 4);
  be32(w + 8, crc.run(0xffffffffu, w + 4, 4) ^ 0xffffffffu);
  return RTCLJ_OK;
}

// ---- P3 reader (the input side of ppm->png): whitespace-separated decimal tokens
namespace {
inline bool p3_ws(char c) { return c == ' ' || c == '\n' || c == '\r' || c == '\t'; }
// true iff `body` holds exactly `n` decimal tokens, each <= maxv, separated by white space, and they were written
// to `out`; false (out unspecified) for anything else -- the caller then runs the sequential parser for the error
bool parse_p3_body_parallel(const char* body, size_t len, unsigned maxv, uint8_t* out, size_t n) {
  unsigned hw = std::thread::hardware_concurrency();
  const unsigned T = std::max(1u, std::min(hw ? hw : 1u, 16u));
  if (T < 2) return false;
  std::vector<size_t> cut(T + 1);
  cut[0] = 0; cut[T] = len;
  for (unsigned k = 1; k < T; ++k) {  // move each cut forward to the start of a token (or of white space)
    size_t c = std::max(cut[k - 1], len / T * k);
    while (c < len && c > 0 && !p3_ws(body[c - 1])) ++c;
    cut[k] = c;
  }
  std::vector<std::vector<uint8_t>> part(T);
  std::vector<char> ok(T, 1);
  auto work = [&](unsigned k) {
    // a private buffer and a private cursor: the vectors of `part` sit side by side in memory, and growing them
    // in place from sixteen threads made every push a cache-line ping-pong (measured: 2.8x slower per thread)
    std::vector<uint8_t> v((cut[k + 1] - cut[k]) / 2 + 16);
    uint8_t* w = v.data();
    const char* p = body + cut[k];
    const char* const e = body + cut[k + 1];
    bool good = true;
    while (p < e) {
      const char c = *p;
      if (p3_ws(c)) { ++p; continue; }
      if (c < '0' || c > '9') { good = false; break; }
      unsigned val = 0, digits = 0;
      while (p < e && *p >= '0' && *p <= '9') { val = val * 10u + (unsigned)(*p++ - '0'); if (++digits > 9) break; }
      if (digits > 9 || val > maxv || (p < e && !p3_ws(*p))) { good = false; break; }
      *w++ = (uint8_t)val;   // (a token needs >= 2 bytes of text except the stretch's last: the buffer cannot overflow)
    }
    v.resize((size_t)(w - v.data()));
    part[k] = std::move(v);
    ok[k] = good ? 1 : 0;
  };
  std::vector<std::thread> threads;
  try {
    for (unsigned k = 1; k < T; ++k) threads.emplace_back(work, k);
    work(0);
  } catch (...) {
    for (auto& t : threads) t.join();
    return false;
  }
  for (auto& t : threads) t.join();
  size_t total = 0;
  for (unsigned k = 0; k < T; ++k) { if (!ok[k]) return false; total += part[k].size(); }
  if (total != n) return false;
  size_t off = 0;
  for (unsigned k = 0; k < T; ++k) { if (!part[k].empty()) std::memcpy(out + off, part[k].data(), part[k].size()); off += part[k].size(); }
  return true;
}
}  // namespace

int rtclj_decode_ppm_p3(const char* text, size_t len, int32_t* width, int32_t* height, uint8_t* out_rgb8,
                        size_t capacity) {
  if (!text || !width || !height) return rtclj_fail(RTCLJ_E_INVALID, "rtclj_decode_ppm_p3: null argument");
  size_t pos = 0;
  auto skip_ws = [&]() { while (pos < len && (text[pos] == ' ' || text[pos] == '\n' || text[pos] == '\r' || text[pos] == '\t')) ++pos; };
  auto number = [&](long long& v) -> bool {  // one non-negative decimal token
    skip_ws();
    if (pos >= len || text[pos] < '0' || text[pos] > '9') return false;
    v = 0;
    while (pos < len && text[pos] >= '0' && text[pos] <= '9') {
      v = v * 10 + (text[pos++] - '0');
      if (v > 0x7fffffffLL) return false;
    }
    return pos >= len || text[pos] == ' ' || text[pos] == '\n' || text[pos] == '\r' || text[pos] == '\t';
  };
  skip_ws();
  if (pos + 2 > len || text[pos] != 'P' || text[pos + 1] != '3') return rtclj_fail(RTCLJ_E_INVALID, "not a P3 file (bad magic)");
  pos += 2;
  long long w = 0, h = 0, maxv = 0;
  if (!number(w) || !number(h) || !number(maxv)) return rtclj_fail(RTCLJ_E_INVALID, "P3 header: expected width, height, maximum");
  if (w <= 0 || h <= 0 || maxv > 255) return rtclj_fail(RTCLJ_E_INVALID, "P3 header: bad dimensions or maximum > 255");
  *width = (int32_t)w;
  *height = (int32_t)h;
  if (!out_rgb8) return RTCLJ_OK;
  const size_t n = (size_t)w * (size_t)h * 3;
  if (capacity < n) return rtclj_fail(RTCLJ_E_BUFFER, "decoded image needs %zu bytes, capacity is %zu", (size_t)n, capacity);
  // Large bodies (the 99 MB of a 3840x2160 scene.ppm) are parsed by several threads, each on a stretch of the text
  // cut at whitespace; anything but a clean body of exactly W*H*3 values in range falls through to the
  // sequential parser below, which reports the error exactly as before.
  if (len - pos >= ((size_t)4 << 20) && parse_p3_body_parallel(text + pos, len - pos, (unsigned)maxv, out_rgb8, n)) return RTCLJ_OK;
  for (size_t i = 0; i < n; ++i) {
    long long v = 0;
    if (!number(v) || v > maxv) return rtclj_fail(RTCLJ_E_INVALID, "P3 body: missing value or value above the maximum");
    out_rgb8[i] = (uint8_t)v;
  }
  skip_ws();
  return pos == len ? RTCLJ_OK : RTCLJ_E_INVALID;  // trailing garbage / more values than W*H
}

// clojure.lang.Ratio.doubleValue: BigDecimal(num).divide(BigDecimal(den), DECIMAL64).doubleValue(),
// i.e. the quotient rounded HALF_EVEN to 16 significant decimal digits, then to double
// (SURVEY.md Appendix B.1).  Integral quotients are Longs in Clojure and stay exact.
double rtclj_ratio_to_double(int64_t num, int64_t den) {
  if (den == 0) return num == 0 ? NAN : (num > 0 ? INFINITY : -INFINITY);
  const int64_t g = gcd64(num, den);
  if (g) { num /= g; den /= g; }
  if (den < 0) { num = -num; den = -den; }
  if (den == 1) return (double)num;
  const bool negative = num < 0;
  unsigned __int128 n = (unsigned __int128)(negative ? -(__int128)num : (__int128)num);
  const unsigned __int128 d = (unsigned __int128)den;
  unsigned __int128 ip = n / d, rem = n % d;
  std::string digits;  // significant digits, no leading zeros
  int exp10 = 0;       // value = 0.DIGITS * 10^exp10
  if (ip > 0) {
    std::string s;
    while (ip > 0) { s.insert(s.begin(), (char)('0' + (int)(ip % 10))); ip /= 10; }
    digits = s;
    exp10 = (int)s.size();
  }
  while ((int)digits.size() < 17 && (rem != 0 || !digits.empty())) {
    rem *= 10;
    const int q = (int)(rem / d);
    rem %= d;
    if (digits.empty() && q == 0) { exp10--; continue; }
    digits.push_back((char)('0' + q));
    if (rem == 0 && (int)digits.size() >= 17) break;
    if (rem == 0) break;
  }
  if ((int)digits.size() > 16) {  // round half-even at 16 digits
    const int guard = digits[16] - '0';
    const bool sticky = rem != 0;
    digits.resize(16);
    bool up = guard > 5 || (guard == 5 && (sticky || ((digits[15] - '0') & 1)));
    if (up) {
      int i = 15;
      while (i >= 0 && digits[(size_t)i] == '9') { digits[(size_t)i] = '0'; --i; }
      if (i >= 0) digits[(size_t)i]++;
      else { digits.insert(digits.begin(), '1'); digits.resize(16); exp10++; }
    }
  }
  const std::string text = std::string(negative ? "-" : "") + "0." + digits + "e" + std::to_string(exp10);
  return std::strtod(text.c_str(), nullptr);
}

// raytracing.clj:105-139
int rtclj_camera_main(int32_t width, int32_t height, double vfov, const double look_from[3],
                      const double look_at[3], const double vup[3], double defocus_angle, double focus_dist,
                      rtclj_camera* out) {
  if (!out || !look_from || !look_at || !vup || width <= 0 || height <= 0) return rtclj_fail(RTCLJ_E_INVALID, "camera: null argument or non-positive size");
  const double theta = deg_to_rad(vfov);
  const double h = std::tan(theta / 2);
  const double viewport_height = 2.0 * h * focus_dist;
  const double viewport_width = viewport_height * rtclj_ratio_to_double(width, height);
  const V w = unit(sub(ld(look_from), ld(look_at)));
  const V u = unit(cross(ld(vup), w));
  const V v = cross(w, u);
  const V center = ld(look_from);
  const V viewport_u = muls(u, viewport_width);
  const V viewport_v = muls(neg(v), viewport_height);
  const V du = divs(viewport_u, (double)width);
  const V dv = divs(viewport_v, (double)height);
  const V upper_left = sub(sub(sub(center, muls(w, focus_dist)), divs(viewport_u, 2.0)), divs(viewport_v, 2.0));
  const V p00 = add(upper_left, muls(add(du, dv), 0.5));
  const double defocus_radius = focus_dist * std::tan(deg_to_rad(defocus_angle / 2.0));
  st(out->pixel00, p00); st(out->pixel_du, du); st(out->pixel_dv, dv); st(out->center, center);
  st(out->defocus_u, muls(u, defocus_radius)); st(out->defocus_v, muls(v, defocus_radius));
  out->defocus_angle = defocus_angle;
  out->width = width; out->height = height;
  return RTCLJ_OK;
}

// realm/raytracing.clj:264-280, 306-322
int rtclj_camera_realm(int32_t width, int32_t height, double vfov, const double look_from[3],
                       const double look_at[3], const double vup[3], rtclj_camera* out) {
  if (!out || !look_from || !look_at || !vup || width <= 0 || height <= 0) return rtclj_fail(RTCLJ_E_INVALID, "camera: null argument or non-positive size");
  const V temp = sub(ld(look_from), ld(look_at));
  const double focal = length(temp);
  const V w = divs(temp, focal);
  const V u = unit(cross(ld(vup), w));
  const V v = cross(w, u);
  const double theta = deg_to_rad(vfov);
  const double h = std::tan(theta / 2.0);
  const double viewport_height = 2.0 * h * focal;
  const double viewport_width = viewport_height * ((double)width / (double)height);
  const V viewport_u = muls(u, viewport_width);
  const V viewport_v = muls(v, -viewport_height);
  const V du = divs(viewport_u, (double)width);
  const V dv = divs(viewport_v, (double)height);
  V ul = sub(ld(look_from), muls(w, focal));
  ul = sub(ul, divs(viewport_u, 2.0));
  ul = sub(ul, divs(viewport_v, 2.0));
  st(out->pixel00, add(ul, divs(add(du, dv), 2.0)));
  st(out->pixel_du, du); st(out->pixel_dv, dv); st(out->center, ld(look_from));
  st(out->defocus_u, mk(0, 0, 0)); st(out->defocus_v, mk(0, 0, 0));
  out->defocus_angle = 0.0;
  out->width = width; out->height = height;
  return RTCLJ_OK;
}

// experimental/raytracing_i.clj:82-90, 127-144
int rtclj_camera_i(int32_t width, int32_t height, rtclj_camera* out) {
  if (!out || width <= 0 || height <= 0) return rtclj_fail(RTCLJ_E_INVALID, "camera: null out or non-positive size");
  const double focal_length = 1.0, viewport_height = 2.0;
  const double viewport_width = viewport_height * rtclj_ratio_to_double(width, height);
  const V center = mk(0, 0, 0);
  const V viewport_u = mk(viewport_width, 0.0, 0.0);
  const V viewport_v = mk(0.0, -viewport_height, 0.0);
  const V du = divs(viewport_u, (double)width);
  const V dv = divs(viewport_v, (double)height);
  V ul = sub(center, mk(0.0, 0.0, focal_length));
  ul = sub(ul, divs(viewport_u, 2.0));
  ul = sub(ul, divs(viewport_v, 2.0));
  st(out->pixel00, add(ul, divs(add(du, dv), 2.0)));
  st(out->pixel_du, du); st(out->pixel_dv, dv); st(out->center, center);
  st(out->defocus_u, mk(0, 0, 0)); st(out->defocus_v, mk(0, 0, 0));
  out->defocus_angle = 0.0;
  out->width = width; out->height = height;
  return RTCLJ_OK;
}

// The book's random-sphere field; same draws as scenes.py.
int rtclj_scene_random_field(uint64_t seed, int32_t lo, int32_t hi, int32_t cap, double* center_xyz,
                             double* radius, int32_t* material, double* albedo_rgb, double* fuzz, double* ior,
                             int32_t* n_out) {
  if (!n_out || hi < lo) return rtclj_fail(RTCLJ_E_INVALID, "scene_random_field: null n_out or hi < lo");
  const bool write = cap > 0;
  if (write && (!center_xyz || !radius || !material || !albedo_rgb || !fuzz || !ior)) return rtclj_fail(RTCLJ_E_INVALID, "scene_random_field: null output array");
  int n = 0;
  bool overflow = false;
  auto put = [&](double cx, double cy, double cz, double r, int kind, double ar, double ag, double ab, double fz,
                 double ri) {
    if (write) {
      if (n >= cap) { overflow = true; ++n; return; }
      center_xyz[3 * n] = cx; center_xyz[3 * n + 1] = cy; center_xyz[3 * n + 2] = cz;
      radius[n] = r; material[n] = kind;
      albedo_rgb[3 * n] = ar; albedo_rgb[3 * n + 1] = ag; albedo_rgb[3 * n + 2] = ab;
      fuzz[n] = fz; ior[n] = ri;
    }
    ++n;
  };
  SplitMix64 rng{seed};
  put(0.0, -1000.0, 0.0, 1000.0, RTCLJ_LAMBERTIAN, 0.5, 0.5, 0.5, 0.0, 1.0);
  for (int a = lo; a < hi; ++a) {
    for (int b = lo; b < hi; ++b) {
      const double choose = rng.uniform();
      const double cx = (double)a + 0.9 * rng.uniform();
      const double cz = (double)b + 0.9 * rng.uniform();
      double d[7];
      for (double& x : d) x = rng.uniform();
      const double dx = cx - 4.0, dz = cz - 0.0;
      if (std::sqrt(dx * dx + dz * dz) <= 0.9) continue;
      if (choose < 0.8) put(cx, 0.2, cz, 0.2, RTCLJ_LAMBERTIAN, d[0] * d[1], d[2] * d[3], d[4] * d[5], 0.0, 1.0);
      else if (choose < 0.95)
        put(cx, 0.2, cz, 0.2, RTCLJ_METAL, 0.5 + 0.5 * d[0], 0.5 + 0.5 * d[1], 0.5 + 0.5 * d[2], 0.5 * d[3], 1.0);
      else put(cx, 0.2, cz, 0.2, RTCLJ_DIELECTRIC, 1.0, 1.0, 1.0, 0.0, 1.5);
    }
  }
  put(0.0, 1.0, 0.0, 1.0, RTCLJ_DIELECTRIC, 1.0, 1.0, 1.0, 0.0, 1.5);
  put(-4.0, 1.0, 0.0, 1.0, RTCLJ_LAMBERTIAN, 0.4, 0.2, 0.1, 0.0, 1.0);
  put(4.0, 1.0, 0.0, 1.0, RTCLJ_METAL, 0.7, 0.6, 0.5, 0.0, 1.0);
  *n_out = n;
  return overflow ? RTCLJ_E_BUFFER : RTCLJ_OK;
}

}  // extern "C"
