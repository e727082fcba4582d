// tile_reader.cpp -- native reader of the reference's per-sample .npz archives (include/mau_tiles.h).
//
// One sample = one ZIP written by np.savez_compressed (reference src/data/processing_10m/process.py:187)
// with the members input.npy [23,H,W], target.npy [2,H,W], metadata.npy [4], temperature_serie.npy [T].
// The reference reads it with np.load on the training thread (src/dataset.py:54-59) and stacks / pads the
// batch in collate_fn (src/dataset.py:87-108).  Here a batch is decoded by a pool of worker threads straight
// into the caller's (pinned) batch buffers: the ZIP central directory (one pread of the file's tail) gives the
// member extents, each member is inflated directly into its slot of the batch (NPY header peeled off the
// front of the same deflate stream) or, if stored, pread there; CRC-32 checked like zipfile does and rows flipped
// while the decoded bytes are still in cache.
// Host-only: C++17, pthreads, zlib.
#include "../../include/mau_tiles.h"
#include "inflate_fast.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

struct Fail {
  int code;
  std::string msg;
};

[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Fail{code, buf};
}

inline uint16_t rd16(const uint8_t* p) { return uint16_t(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
inline uint64_t rd64(const uint8_t* p) { return uint64_t(rd32(p)) | (uint64_t(rd32(p + 4)) << 32); }

// ---- CRC-32 (the ZIP / zlib polynomial 0xEDB88320) ---------------------------------------------------------
// zipfile verifies every member it reads; with the decoder above no longer the bottleneck zlib's table-driven
// crc32() (~1.8 GB/s) was a quarter of a sample's decode time.  On x86-64 with PCLMULQDQ the bulk is folded 64
// bytes at a time with carry-less multiplies (Gopal et al., "Fast CRC Computation for Generic Polynomials Using
// PCLMULQDQ Instruction", Intel 2009: fold-by-4, fold to 128 bits, 128 -> 64 -> 32 with a Barrett reduction);
// heads, tails and other CPUs go through zlib.  Checked against zlib in tests/test_tiles_cpu.py.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) uint32_t crc32_clmul(const uint8_t* p, size_t n /* multiple of 16, >= 64 */, uint32_t raw) {
  // x^(512+32), x^(512-32); x^(128+32), x^(128-32); x^64; P(x), floor(x^64 / P(x)) -- bit-reflected
  const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596, 0x0154442bd4);
  const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009e, 0x01751997d0);
  const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124);
  const __m128i poly = _mm_set_epi64x(0x01f7011641, 0x01db710641);
#define MAU_LD(q) _mm_loadu_si128(reinterpret_cast<const __m128i*>(q))
#define MAU_FOLD(x, k, next) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x, k, 0x00), _mm_clmulepi64_si128(x, k, 0x11)), next)
  __m128i a = _mm_xor_si128(MAU_LD(p), _mm_cvtsi32_si128(int(raw))), b = MAU_LD(p + 16), c = MAU_LD(p + 32), d = MAU_LD(p + 48);
  p += 64;
  n -= 64;
  while (n >= 64) {
    a = MAU_FOLD(a, k1k2, MAU_LD(p));
    b = MAU_FOLD(b, k1k2, MAU_LD(p + 16));
    c = MAU_FOLD(c, k1k2, MAU_LD(p + 32));
    d = MAU_FOLD(d, k1k2, MAU_LD(p + 48));
    p += 64;
    n -= 64;
  }
  a = MAU_FOLD(a, k3k4, b);
  a = MAU_FOLD(a, k3k4, c);
  a = MAU_FOLD(a, k3k4, d);
  while (n >= 16) {
    a = MAU_FOLD(a, k3k4, MAU_LD(p));
    p += 16;
    n -= 16;
  }
  const __m128i lo32 = _mm_setr_epi32(~0, 0, ~0, 0);
  __m128i t = _mm_clmulepi64_si128(a, k3k4, 0x10);                  // 128 -> 96 bits
  a = _mm_xor_si128(_mm_srli_si128(a, 8), t);
  t = _mm_srli_si128(a, 4);                                         // 96 -> 64 bits
  a = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(a, lo32), k5, 0x00), t);
  t = _mm_clmulepi64_si128(_mm_and_si128(a, lo32), poly, 0x10);     // Barrett: 64 -> 32 bits
  t = _mm_clmulepi64_si128(_mm_and_si128(t, lo32), poly, 0x00);
  return uint32_t(_mm_extract_epi32(_mm_xor_si128(a, t), 1));
#undef MAU_LD
#undef MAU_FOLD
}
bool have_clmul() {
  static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
  return ok;
}
#endif

// running CRC-32 in zlib's convention (finalised value in, finalised value out; start from 0)
uint32_t crc32_update(uint32_t crc, const uint8_t* p, uint64_t n) {
#if defined(__x86_64__)
  if (n >= 64 && have_clmul()) {
    uint64_t bulk = n & ~uint64_t(15);
    crc = ~crc32_clmul(p, size_t(bulk), ~crc);
    p += bulk;
    n -= bulk;
  }
#endif
  while (n) {  // zlib's crc32() takes a 32-bit length
    uInt k = uInt(std::min<uint64_t>(n, 1u << 30));
    crc = uint32_t(crc32(crc, p, k));
    p += k;
    n -= k;
  }
  return crc;
}

// ---- one archive, read with pread ------------------------------------------------------------------------
// The first version mmap'ed every archive in every task.  With 16 decode threads in one process that serialises on
// the address-space lock (mmap / munmap take it exclusively, every munmap ends in TLB-shootdown interrupts on all
// cores running our threads -- including the one launching the GPU step) and costs a page fault per 4 KiB read:
// stored archives decoded at 819 tiles/s with all threads busy against 1 460 when throttled by the training step.
// pread() into the destination (stored fp32 payloads: zero extra copies) or into a per-thread arena (compressed
// members, the directory) has none of that.
std::vector<uint8_t>& arena() {
  thread_local std::vector<uint8_t> a;
  return a;
}

struct Archive {
  int fd = -1;
  size_t size = 0;
  std::string path;
  explicit Archive(const std::string& p) : path(p) {
    fd = ::open(p.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) fail(MAU_TILES_E_IO, "cannot open '%s': %s", p.c_str(), strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
      int e = errno;
      ::close(fd);
      fd = -1;
      fail(MAU_TILES_E_IO, "cannot stat '%s': %s", p.c_str(), strerror(e));
    }
    size = size_t(st.st_size);
    if (size == 0) {
      ::close(fd);
      fd = -1;
      fail(MAU_TILES_E_FORMAT, "'%s' is empty (not a zip archive)", p.c_str());
    }
  }
  ~Archive() {
    if (fd >= 0) ::close(fd);
  }
  Archive(const Archive&) = delete;
  Archive& operator=(const Archive&) = delete;
  void check_range(uint64_t off, uint64_t len) const {
    if (off > size || len > size - off) fail(MAU_TILES_E_FORMAT, "'%s': truncated archive (need %llu bytes at %llu of %zu)", path.c_str(),
                                             (unsigned long long)len, (unsigned long long)off, size);
  }
  // exactly len bytes at off -> dst
  void read_at(void* dst, uint64_t off, uint64_t len) const {
    check_range(off, len);
    uint8_t* d = static_cast<uint8_t*>(dst);
    while (len) {
      ssize_t r = ::pread(fd, d, size_t(std::min<uint64_t>(len, 1u << 30)), off_t(off));
      if (r < 0) {
        if (errno == EINTR) continue;
        fail(MAU_TILES_E_IO, "cannot read '%s': %s", path.c_str(), strerror(errno));
      }
      if (r == 0) fail(MAU_TILES_E_FORMAT, "'%s': truncated archive (file shrank while reading)", path.c_str());
      d += r;
      off += uint64_t(r);
      len -= uint64_t(r);
    }
  }
  // the range copied into this thread's arena, followed by `slack` zero bytes; valid until the next fetch on this thread
  const uint8_t* fetch(uint64_t off, uint64_t len, size_t slack = 0) const {
    check_range(off, len);
    std::vector<uint8_t>& a = arena();
    if (a.size() < len + slack) a.resize(size_t(len + slack));
    read_at(a.data(), off, len);
    if (slack) memset(a.data() + len, 0, slack);
    return a.data();
  }
};

// ---- ZIP central directory -------------------------------------------------------------------------------
struct Member {
  uint16_t method = 0;  // 0 stored, 8 deflate
  uint16_t flags = 0;   // general purpose bits (bit 0: encrypted)
  uint32_t crc = 0;
  uint32_t dos_time_date = 0;
  uint64_t csize = 0, usize = 0;
  uint64_t data_off = 0;  // file offset of the first byte of the (compressed) payload
  bool found = false;
};

enum { M_INPUT = 0, M_TARGET = 1, M_METADATA = 2, M_SERIES = 3, M_COUNT = 4 };
const char* const kMemberNames[M_COUNT] = {"input.npy", "target.npy", "metadata.npy", "temperature_serie.npy"};

// Walks the central directory and calls visit(name, name_len, member) for every entry (payload extent validated).
template <typename Visit>
void walk_zip(const Archive& mp, Visit&& visit) {
  const size_t n = mp.size;
  if (n < 22) fail(MAU_TILES_E_FORMAT, "'%s' is not a zip archive (too short)", mp.path.c_str());
  // the tail of the file holds the end-of-central-directory record (last 22 bytes + up to 64 KiB of comment) and,
  // for the small directories of these archives, the central directory itself: one pread gets both
  thread_local std::vector<uint8_t> tail, spill;
  size_t tail_len = 0, tail_off = 0, eocd = size_t(-1);
  for (size_t want : {size_t(8192), size_t(22 + 65535)}) {
    size_t len = std::min(n, want);
    if (len == tail_len) break;
    tail_len = len;
    tail_off = n - len;
    if (tail.size() < len) tail.resize(len);
    mp.read_at(tail.data(), tail_off, len);
    for (size_t p = len - 22 + 1; p-- > 0;) {
      if (rd32(tail.data() + p) == 0x06054b50u) {
        eocd = tail_off + p;
        break;
      }
    }
    if (eocd != size_t(-1)) break;
  }
  if (eocd == size_t(-1)) fail(MAU_TILES_E_FORMAT, "'%s' is not a zip archive (no end-of-central-directory record)", mp.path.c_str());
  auto view = [&](uint64_t off, uint64_t len) -> const uint8_t* {  // bytes of the file: from the tail if inside, else read
    mp.check_range(off, len);
    if (off >= tail_off) return tail.data() + (off - tail_off);
    if (spill.size() < len) spill.resize(size_t(len));
    mp.read_at(spill.data(), off, len);
    return spill.data();
  };
  const uint8_t* e = tail.data() + (eocd - tail_off);
  uint64_t entries = rd16(e + 10), cd_size = rd32(e + 12), cd_off = rd32(e + 16);
  if (entries == 0xFFFFu || cd_size == 0xFFFFFFFFu || cd_off == 0xFFFFFFFFu) {
    // ZIP64: locator sits right in front of the EOCD record
    if (eocd < 20) fail(MAU_TILES_E_FORMAT, "'%s': zip64 locator missing", mp.path.c_str());
    uint8_t loc[20];
    mp.read_at(loc, eocd - 20, 20);
    if (rd32(loc) != 0x07064b50u) fail(MAU_TILES_E_FORMAT, "'%s': zip64 locator missing", mp.path.c_str());
    uint8_t z[56];
    mp.read_at(z, rd64(loc + 8), 56);
    if (rd32(z) != 0x06064b50u) fail(MAU_TILES_E_FORMAT, "'%s': bad zip64 end record", mp.path.c_str());
    entries = rd64(z + 32);
    cd_size = rd64(z + 40);
    cd_off = rd64(z + 48);
  }
  if (cd_size > (uint64_t(1) << 31)) fail(MAU_TILES_E_FORMAT, "'%s': implausible central directory", mp.path.c_str());
  const uint8_t* cd = view(cd_off, cd_size);
  uint64_t pos = 0;
  for (uint64_t i = 0; i < entries; ++i) {
    if (pos + 46 > cd_size || rd32(cd + pos) != 0x02014b50u) fail(MAU_TILES_E_FORMAT, "'%s': bad central directory entry %llu", mp.path.c_str(), (unsigned long long)i);
    const uint8_t* h = cd + pos;
    uint16_t flags = rd16(h + 8), method = rd16(h + 10);
    uint32_t crc = rd32(h + 16);
    uint64_t csize = rd32(h + 20), usize = rd32(h + 24);
    uint16_t nlen = rd16(h + 28), xlen = rd16(h + 30), clen = rd16(h + 32);
    uint64_t lho = rd32(h + 42);
    if (pos + 46 + uint64_t(nlen) + xlen + clen > cd_size) fail(MAU_TILES_E_FORMAT, "'%s': central directory overruns", mp.path.c_str());
    const char* name = reinterpret_cast<const char*>(h + 46);
    // zip64 extended information: present values in the order usize, csize, local header offset
    const uint8_t* x = h + 46 + nlen;
    for (uint32_t xp = 0; xp + 4 <= xlen;) {
      uint16_t id = rd16(x + xp), sz = rd16(x + xp + 2);
      if (xp + 4u + sz > xlen) break;
      if (id == 0x0001) {
        uint32_t q = xp + 4, qe = xp + 4 + sz;
        if (usize == 0xFFFFFFFFu && q + 8 <= qe) { usize = rd64(x + q); q += 8; }
        if (csize == 0xFFFFFFFFu && q + 8 <= qe) { csize = rd64(x + q); q += 8; }
        if (lho == 0xFFFFFFFFu && q + 8 <= qe) { lho = rd64(x + q); q += 8; }
      }
      xp += 4u + sz;
    }
    {
      uint8_t lh[30];
      mp.read_at(lh, lho, 30);
      if (rd32(lh) != 0x04034b50u) fail(MAU_TILES_E_FORMAT, "'%s': bad local header of %.*s", mp.path.c_str(), int(nlen), name);
      uint64_t start = lho + 30 + rd16(lh + 26) + rd16(lh + 28);
      Member M;
      M.method = method;
      M.flags = flags;
      M.crc = crc;
      M.dos_time_date = rd32(h + 12);
      M.csize = csize;
      M.usize = usize;
      mp.check_range(start, csize);
      M.data_off = start;
      M.found = true;
      visit(name, size_t(nlen), M);
    }
    pos += 46 + uint64_t(nlen) + xlen + clen;
  }
}

void check_supported(const Archive& mp, const Member& M, const char* name, size_t nlen) {
  if (M.flags & 1) fail(MAU_TILES_E_FORMAT, "'%s': member %.*s is encrypted", mp.path.c_str(), int(nlen), name);
  if (M.method != 0 && M.method != 8)
    fail(MAU_TILES_E_FORMAT, "'%s': member %.*s uses compression method %u (only stored/deflate)", mp.path.c_str(), int(nlen), name, M.method);
}

// the four members of a sample (duplicate names: zipfile keeps the last entry, so do we)
void parse_zip(const Archive& mp, Member out[M_COUNT]) {
  walk_zip(mp, [&](const char* name, size_t nlen, const Member& M) {
    for (int m = 0; m < M_COUNT; ++m) {
      if (nlen == strlen(kMemberNames[m]) && memcmp(name, kMemberNames[m], nlen) == 0) {
        check_supported(mp, M, name, nlen);
        out[m] = M;
      }
    }
  });
}

// ---- sequential reader over one member (stored or raw deflate), CRC accumulated on the way ----------------
class MemberStream {
 public:
  // `max_in` bounds how much of a deflated member is fetched (the NPY header sits in the first few hundred bytes:
  // probing a sample must not read its megabytes).  The compressed bytes live in this thread's arena: no other
  // fetch on this thread while the stream is in use.
  MemberStream(const Archive& mp, const Member& m, const char* name, bool check_crc, uint64_t max_in = ~uint64_t(0))
      : mp_(mp), m_(m), name_(name), check_crc_(check_crc) {
    if (m.method == 8) {
      memset(&z_, 0, sizeof z_);
      if (inflateInit2(&z_, -15) != Z_OK) fail(MAU_TILES_E_FORMAT, "zlib inflateInit2 failed");
      z_init_ = true;
      in_left_ = std::min(m.csize, max_in);
      z_.next_in = const_cast<Bytef*>(mp.fetch(m.data_off, in_left_));
      z_.avail_in = 0;
    }
    crc_ = 0;
  }
  ~MemberStream() {
    if (z_init_) inflateEnd(&z_);
  }
  MemberStream(const MemberStream&) = delete;
  MemberStream& operator=(const MemberStream&) = delete;

  void read(void* dst, uint64_t n) {
    if (n > m_.usize - produced_) fail(MAU_TILES_E_FORMAT, "'%s': member %s is shorter than its NPY header says (%llu of %llu bytes)", mp_.path.c_str(), name_,
                                       (unsigned long long)(m_.usize - produced_), (unsigned long long)n);
    uint8_t* out = static_cast<uint8_t*>(dst);
    if (m_.method == 0) {
      mp_.read_at(out, m_.data_off + produced_, n);
    } else {
      uint64_t left = n;
      uint8_t* o = out;
      while (left) {
        if (z_.avail_in == 0 && in_left_) {
          uInt take = uInt(std::min<uint64_t>(in_left_, 1u << 30));
          z_.avail_in = take;
          in_left_ -= take;
        }
        uInt want = uInt(std::min<uint64_t>(left, 1u << 30));
        z_.next_out = o;
        z_.avail_out = want;
        int rc = inflate(&z_, Z_NO_FLUSH);
        uInt got = want - z_.avail_out;
        o += got;
        left -= got;
        if (rc == Z_STREAM_END) {
          if (left) fail(MAU_TILES_E_FORMAT, "'%s': deflate stream of %s ends early", mp_.path.c_str(), name_);
          break;
        }
        if (rc != Z_OK || (got == 0 && z_.avail_in == 0 && in_left_ == 0))
          fail(MAU_TILES_E_FORMAT, "'%s': error while decompressing %s (%s)", mp_.path.c_str(), name_, z_.msg ? z_.msg : "truncated stream");
      }
    }
    if (check_crc_) crc_ = crc32_update(crc_, out, n);
    produced_ += n;
  }
  void finish() {
    if (produced_ != m_.usize) fail(MAU_TILES_E_FORMAT, "'%s': member %s holds %llu bytes, NPY header accounts for %llu", mp_.path.c_str(), name_,
                                    (unsigned long long)m_.usize, (unsigned long long)produced_);
    if (check_crc_ && crc_ != m_.crc) fail(MAU_TILES_E_FORMAT, "'%s': bad CRC-32 for member %s", mp_.path.c_str(), name_);
  }

 private:
  const Archive& mp_;
  const Member& m_;
  const char* name_;
  bool check_crc_;
  z_stream z_;
  bool z_init_ = false;
  uint64_t in_left_ = 0, produced_ = 0;
  uint32_t crc_ = 0;
};

// ---- NPY header ------------------------------------------------------------------------------------------
enum DType { F4, F8, F2, I1, U1, I2, U2, I4, U4, I8, U8, B1 };
struct NpyHeader {
  DType dtype = F4;
  int itemsize = 4;
  int ndim = 0;
  int64_t shape[8] = {0};
  int64_t count = 1;
};

bool find_key(const std::string& h, const char* key, size_t* value_pos) {
  std::string k1 = std::string("'") + key + "'", k2 = std::string("\"") + key + "\"";
  size_t p = h.find(k1);
  if (p == std::string::npos) p = h.find(k2);
  if (p == std::string::npos) return false;
  p = h.find(':', p);
  if (p == std::string::npos) return false;
  ++p;
  while (p < h.size() && (h[p] == ' ' || h[p] == '\t')) ++p;
  *value_pos = p;
  return true;
}

// magic + version + header length: returns the dict length, *prefix = 10 (format 1.0) or 12 (2.0 / 3.0)
uint32_t npy_prefix(const uint8_t* pre, size_t avail, size_t* prefix, const Archive& mp, const char* name) {
  if (avail < 10 || memcmp(pre, "\x93NUMPY", 6) != 0)
    fail(MAU_TILES_E_FORMAT, "'%s': member %s is not an NPY array (object arrays / pickles are not supported)", mp.path.c_str(), name);
  uint32_t hlen;
  if (pre[6] == 1) {
    hlen = rd16(pre + 8);
    *prefix = 10;
  } else if (pre[6] == 2 || pre[6] == 3) {
    if (avail < 12) fail(MAU_TILES_E_FORMAT, "'%s': member %s: truncated NPY header", mp.path.c_str(), name);
    hlen = rd32(pre + 8);
    *prefix = 12;
  } else {
    fail(MAU_TILES_E_FORMAT, "'%s': member %s has NPY format version %u.%u", mp.path.c_str(), name, pre[6], pre[7]);
  }
  if (hlen > (1u << 20)) fail(MAU_TILES_E_FORMAT, "'%s': member %s has an implausible NPY header (%u bytes)", mp.path.c_str(), name, hlen);
  return hlen;
}

NpyHeader parse_npy_dict(const std::string& h, const Archive& mp, const char* name);

NpyHeader read_npy_header(MemberStream& s, const Archive& mp, const char* name) {
  uint8_t pre[12];
  s.read(pre, 10);
  if (pre[6] == 2 || pre[6] == 3) s.read(pre + 10, 2);
  size_t prefix;
  uint32_t hlen = npy_prefix(pre, 12, &prefix, mp, name);
  std::string h(hlen, '\0');
  s.read(h.data(), hlen);
  return parse_npy_dict(h, mp, name);
}

// header of an NPY image held in memory; *header_total = bytes in front of the payload
NpyHeader read_npy_header_mem(const uint8_t* p, size_t avail, size_t* header_total, const Archive& mp, const char* name) {
  size_t prefix;
  uint32_t hlen = npy_prefix(p, avail, &prefix, mp, name);
  if (prefix + hlen > avail) fail(MAU_TILES_E_FORMAT, "'%s': member %s: truncated NPY header", mp.path.c_str(), name);
  *header_total = prefix + hlen;
  return parse_npy_dict(std::string(reinterpret_cast<const char*>(p) + prefix, hlen), mp, name);
}

NpyHeader parse_npy_dict(const std::string& h, const Archive& mp, const char* name) {
  NpyHeader r;
  size_t p;
  if (!find_key(h, "descr", &p) || p >= h.size() || (h[p] != '\'' && h[p] != '"')) fail(MAU_TILES_E_DTYPE, "'%s': member %s: structured or missing dtype", mp.path.c_str(), name);
  size_t q = h.find(h[p], p + 1);
  if (q == std::string::npos) fail(MAU_TILES_E_FORMAT, "'%s': member %s: malformed descr", mp.path.c_str(), name);
  std::string d = h.substr(p + 1, q - p - 1);
  if (d.size() < 3) fail(MAU_TILES_E_DTYPE, "'%s': member %s has dtype '%s'", mp.path.c_str(), name, d.c_str());
  char order = d[0];
  std::string code = d.substr(1);
  struct { const char* c; DType t; int sz; } table[] = {{"f4", F4, 4}, {"f8", F8, 8}, {"f2", F2, 2}, {"i1", I1, 1}, {"u1", U1, 1}, {"i2", I2, 2},
                                                          {"u2", U2, 2}, {"i4", I4, 4}, {"u4", U4, 4}, {"i8", I8, 8}, {"u8", U8, 8}, {"b1", B1, 1}};
  bool ok = false;
  for (auto& t : table)
    if (code == t.c) {
      r.dtype = t.t;
      r.itemsize = t.sz;
      ok = true;
    }
  if (!ok || !(order == '<' || order == '|' || (order == '=') )) fail(MAU_TILES_E_DTYPE, "'%s': member %s has dtype '%s' (little-endian numeric types only)", mp.path.c_str(), name, d.c_str());
  if (!find_key(h, "shape", &p) || p >= h.size() || h[p] != '(') fail(MAU_TILES_E_FORMAT, "'%s': member %s: malformed shape", mp.path.c_str(), name);
  ++p;
  while (p < h.size() && h[p] != ')') {
    if (h[p] == ' ' || h[p] == ',') {
      ++p;
      continue;
    }
    if (h[p] < '0' || h[p] > '9' || r.ndim >= 8) fail(MAU_TILES_E_FORMAT, "'%s': member %s: malformed shape", mp.path.c_str(), name);
    int64_t v = 0;
    while (p < h.size() && h[p] >= '0' && h[p] <= '9') {
      if (v > (int64_t(1) << 56)) fail(MAU_TILES_E_FORMAT, "'%s': member %s: shape overflow", mp.path.c_str(), name);
      v = v * 10 + (h[p++] - '0');
    }
    if (p < h.size() && h[p] == 'L') ++p;  // Python 2 longs
    r.shape[r.ndim++] = v;
  }
  if (p >= h.size()) fail(MAU_TILES_E_FORMAT, "'%s': member %s: malformed shape", mp.path.c_str(), name);
  for (int i = 0; i < r.ndim; ++i) {
    if (r.shape[i] && r.count > (int64_t(1) << 46) / std::max<int64_t>(r.shape[i], 1)) fail(MAU_TILES_E_FORMAT, "'%s': member %s: implausible shape", mp.path.c_str(), name);
    r.count *= r.shape[i];
  }
  if (find_key(h, "fortran_order", &p) && h.compare(p, 4, "True") == 0 && r.ndim > 1)
    fail(MAU_TILES_E_DTYPE, "'%s': member %s is Fortran-ordered", mp.path.c_str(), name);
  return r;
}

inline float half_to_float(uint16_t h) {
  uint32_t s = uint32_t(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u, bits;
  if (e == 0) {
    if (m == 0) {
      bits = s;
    } else {  // subnormal
      int sh = 0;
      while (!(m & 1024u)) {
        m <<= 1;
        ++sh;
      }
      bits = s | ((113u - sh) << 23) | ((m & 1023u) << 13);
    }
  } else if (e == 31) {
    bits = s | 0x7F800000u | (m << 13);
  } else {
    bits = s | ((e + 112u) << 23) | (m << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

template <typename T>
void convert_run(const uint8_t* src, float* dst, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    T v;
    memcpy(&v, src + i * sizeof(T), sizeof(T));
    dst[i] = float(v);  // same rounding as Tensor.float(): round to nearest even
  }
}

void convert(DType t, const uint8_t* src, float* dst, int64_t n) {
  switch (t) {
    case F4: memcpy(dst, src, size_t(n) * 4); break;
    case F8: convert_run<double>(src, dst, n); break;
    case F2:
      for (int64_t i = 0; i < n; ++i) dst[i] = half_to_float(rd16(src + 2 * i));
      break;
    case I1: convert_run<int8_t>(src, dst, n); break;
    case U1: convert_run<uint8_t>(src, dst, n); break;
    case B1:
      for (int64_t i = 0; i < n; ++i) dst[i] = src[i] ? 1.f : 0.f;
      break;
    case I2: convert_run<int16_t>(src, dst, n); break;
    case U2: convert_run<uint16_t>(src, dst, n); break;
    case I4: convert_run<int32_t>(src, dst, n); break;
    case U4: convert_run<uint32_t>(src, dst, n); break;
    case I8: convert_run<int64_t>(src, dst, n); break;
    case U8: convert_run<uint64_t>(src, dst, n); break;
  }
}

// payload of `count` elements -> fp32 at dst (direct inflate for f4, chunked conversion otherwise)
void read_payload(MemberStream& s, const NpyHeader& h, float* dst) {
  if (h.count == 0) return;
  if (h.dtype == F4) {
    s.read(dst, uint64_t(h.count) * 4);
    return;
  }
  const int64_t chunk = 1 << 16;
  std::vector<uint8_t> tmp(size_t(chunk) * h.itemsize);
  for (int64_t done = 0; done < h.count;) {
    int64_t n = std::min(chunk, h.count - done);
    s.read(tmp.data(), uint64_t(n) * h.itemsize);
    convert(h.dtype, tmp.data(), dst + done, n);
    done += n;
  }
}

void reverse_rows(float* p, int64_t rows, int64_t w) {  // np.flip(x, axis=2) of a [C,H,W] array, in place
  for (int64_t r = 0; r < rows; ++r) std::reverse(p + r * w, p + (r + 1) * w);
}

// Decodes one member: NPY header -> `place(header)` validates the shape and names the destination -> payload as
// fp32 there.  Deflated fp32 members (the reference's format) go through the one-shot decoder of inflate_fast.h:
// the first 64 KiB are decoded into a scratch buffer (header + the start of the payload), the rest straight into
// the destination, whose first 64 KiB minus header then serve as the match history.  Stored members are copied
// from the mapping.  Anything else (other dtypes under deflate, MAU_TILES_FLAG_ZLIB) streams through zlib.
// `flip_w` > 0 additionally reverses every run of flip_w floats of the payload (np.flip over the last axis of a
// C-ordered array whose last extent is flip_w).  On the fast path this happens chunk by chunk while the decoded bytes
// are still in cache, as does the CRC: rows are reversed once they lie more than 32 KiB (the DEFLATE window) behind
// the decode position, so the match history in front of the decoder stays in stream order.
template <typename Place>
NpyHeader read_member(const Archive& mp, const Member& M, const char* name, bool check_crc, bool force_zlib, int64_t flip_w, Place&& place) {
  constexpr size_t kHead = 65536, kChunk = 256 * 1024, kWindow = 32768;
  auto check_size = [&](const NpyHeader& h, size_t header_total) {
    if (uint64_t(h.count) * h.itemsize + header_total != M.usize)
      fail(MAU_TILES_E_FORMAT, "'%s': member %s holds %llu bytes, NPY header accounts for %llu", mp.path.c_str(), name, (unsigned long long)M.usize,
           (unsigned long long)(uint64_t(h.count) * h.itemsize + header_total));
  };
  if (M.method == 0) {
    if (M.csize != M.usize) fail(MAU_TILES_E_FORMAT, "'%s': stored member %s has differing sizes", mp.path.c_str(), name);
    uint8_t hb[4096];
    const uint8_t* hp = hb;
    size_t hn = size_t(std::min<uint64_t>(M.usize, sizeof hb)), prefix;
    mp.read_at(hb, M.data_off, hn);
    const uint64_t hdr = uint64_t(npy_prefix(hb, hn, &prefix, mp, name)) + prefix;
    if (hdr > hn) {  // an unusually long header dict
      if (hdr > M.usize) fail(MAU_TILES_E_FORMAT, "'%s': member %s: truncated NPY header", mp.path.c_str(), name);
      hp = mp.fetch(M.data_off, hdr);
      hn = size_t(hdr);
    }
    size_t header_total;
    NpyHeader h = read_npy_header_mem(hp, hn, &header_total, mp, name);
    check_size(h, header_total);
    float* dst = place(h);
    uint32_t crc = check_crc ? crc32_update(0, hp, header_total) : 0;
    const uint64_t payload = uint64_t(h.count) * h.itemsize, at = M.data_off + header_total;
    if (h.dtype == F4) {
      // straight from the page cache into the batch slot, a chunk of whole rows at a time; CRC and flip while in cache
      const uint64_t row_bytes = flip_w > 0 ? uint64_t(flip_w) * 4 : 0;
      const uint64_t step = row_bytes ? std::max<uint64_t>(row_bytes, kChunk / row_bytes * row_bytes) : kChunk;
      uint8_t* d = reinterpret_cast<uint8_t*>(dst);
      for (uint64_t pos = 0; pos < payload; pos += step) {
        const uint64_t nb = std::min(step, payload - pos);
        mp.read_at(d + pos, at + pos, nb);
        if (check_crc) crc = crc32_update(crc, d + pos, nb);
        if (row_bytes)
          for (uint64_t r = 0; r + row_bytes <= nb; r += row_bytes) {
            float* row = reinterpret_cast<float*>(d + pos + r);
            std::reverse(row, row + flip_w);
          }
      }
    } else {
      const uint8_t* src = mp.fetch(at, payload);
      if (check_crc) crc = crc32_update(crc, src, payload);
      convert(h.dtype, src, dst, h.count);
      if (flip_w > 0) reverse_rows(dst, h.count / flip_w, flip_w);
    }
    if (check_crc && crc != M.crc) fail(MAU_TILES_E_FORMAT, "'%s': bad CRC-32 for member %s", mp.path.c_str(), name);
    return h;
  }
  if (!force_zlib) {
    thread_local std::vector<uint8_t> scratch(kHead + 16);
    thread_local mau_inflate::Inflater inf;
    try {
      const size_t c0 = size_t(std::min<uint64_t>(M.usize, kHead));
      uint8_t* S = scratch.data();
      const uint8_t* in = mp.fetch(M.data_off, M.csize, 16);  // + the 16 bytes of slack the decoder's 8-byte loads may touch
      inf.init(in, M.csize, in + M.csize + 16);
      if (inf.run(S, S, S + c0) != S + c0) fail(MAU_TILES_E_FORMAT, "'%s': deflate stream of %s ends early", mp.path.c_str(), name);
      size_t header_total;
      NpyHeader h = read_npy_header_mem(S, c0, &header_total, mp, name);
      if (h.dtype == F4) {
        check_size(h, header_total);
        uint8_t* dst = reinterpret_cast<uint8_t*>(place(h));
        const uint64_t payload = uint64_t(h.count) * 4;
        const size_t n0 = c0 - header_total;
        memcpy(dst, S + header_total, n0);
        uint32_t crc = check_crc ? crc32_update(0, S, c0) : 0;
        const uint64_t row_bytes = uint64_t(flip_w) * 4;
        uint64_t rows_done = 0;
        auto flip_rows_before = [&](uint64_t limit) {  // rows that end at or before byte `limit` of the payload
          if (flip_w <= 0) return;
          for (uint64_t rows = limit / row_bytes; rows_done < rows; ++rows_done) {
            float* r = reinterpret_cast<float*>(dst + rows_done * row_bytes);
            std::reverse(r, r + flip_w);
          }
        };
        for (uint64_t pos = n0; pos < payload;) {
          const uint64_t stop = std::min<uint64_t>(payload, pos + kChunk);
          if (inf.run(dst, dst + pos, dst + stop) != dst + stop) fail(MAU_TILES_E_FORMAT, "'%s': deflate stream of %s ends early", mp.path.c_str(), name);
          if (check_crc) crc = crc32_update(crc, dst + pos, stop - pos);
          pos = stop;
          if (pos > kWindow) flip_rows_before(pos - kWindow);
        }
        if (!inf.done()) fail(MAU_TILES_E_FORMAT, "'%s': member %s holds more data than its directory entry says", mp.path.c_str(), name);
        if (check_crc && crc != M.crc) fail(MAU_TILES_E_FORMAT, "'%s': bad CRC-32 for member %s", mp.path.c_str(), name);
        flip_rows_before(payload);
        return h;
      }
    } catch (const mau_inflate::Error& e) {
      fail(MAU_TILES_E_FORMAT, "'%s': error while decompressing %s (%s)", mp.path.c_str(), name, e.what);
    }
  }
  MemberStream s(mp, M, name, check_crc);
  NpyHeader h = read_npy_header(s, mp, name);
  float* dst = place(h);
  read_payload(s, h, dst);
  s.finish();
  if (flip_w > 0) reverse_rows(dst, h.count / flip_w, flip_w);
  return h;
}

// ---- the worker pool -------------------------------------------------------------------------------------
class Pool {
 public:
  // `nice` > 0 lowers the workers' priority (MAU_TILES_FLAG_NICE)
  Pool(int n, int nice) : nice_(nice) {
    for (int i = 0; i < n; ++i) th_.emplace_back([this] { run(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // `urgent` tasks (header-only reads a caller blocks on) go ahead of the queued decodes
  void push(std::function<void()> f, bool urgent = false) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (urgent) q_.push_front(std::move(f));
      else q_.push_back(std::move(f));
    }
    cv_.notify_one();
  }
  int size() const { return int(th_.size()); }

 private:
  void run() {
    if (nice_ > 0) setpriority(PRIO_PROCESS, id_t(syscall(SYS_gettid)), nice_);  // per-thread on Linux; failure is harmless
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;  // stop requested and drained
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
    }
  }
  std::vector<std::thread> th_;
  std::deque<std::function<void()>> q_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
  int nice_ = 0;
};

struct Batch {
  std::vector<int64_t> idx;
  std::vector<uint8_t> flip;
  int64_t dims[8];
  float *input, *target, *metadata, *series;
  uint16_t* input_staged = nullptr;   // optional: the input tiles once more as bf16 NHWC [n, H, W, staged_cs] (engine.stage_maps)
  int64_t staged_cs = 0;
  int64_t series_stride;
  int64_t* series_len;
  std::mutex mu;
  std::condition_variable cv;
  int64_t pending = 0;
  int code = 0;
  std::string msg;
};

}  // namespace

struct mau_tiles {
  std::vector<std::string> paths;
  int flags = 0;
  std::unique_ptr<Pool> pool;
  std::mutex mu;
  std::map<int64_t, std::shared_ptr<Batch>> inflight;
  int64_t next_ticket = 1;
  std::atomic<int64_t> payload_bytes{0}, archive_bytes{0}, samples{0};
};

namespace {

// which members a task decodes: the `input` member is ~90 % of a sample, so it is its own task
enum { PART_INPUT = 1, PART_REST = 2 };

// fp32 -> bf16, round to nearest even (what torch's .to(bfloat16) and the engine's layout kernel do); NaN -> 0x7FC0
inline uint16_t to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7FC0;
  u += 0x7fffu + ((u >> 16) & 1u);
  return uint16_t(u >> 16);
}

// one decoded tile, planar fp32 [C][H][W] -> the engine's input layout bf16 [H][W][cs] (pad channels zero), on the thread
// that has just decoded it (the planes are still in its cache): the loader then ships 2 * cs instead of 4 * C bytes per pixel
void stage_tile(const float* src, int64_t Cc, int64_t H, int64_t W, uint16_t* dst, int64_t cs) {
  const int64_t HW = H * W;
  for (int64_t h = 0; h < H; ++h) {
    const float* row = src + h * W;
    uint16_t* out = dst + h * W * cs;
    for (int64_t w = 0; w < W; ++w, out += cs) {
      for (int64_t c = 0; c < Cc; ++c) out[c] = to_bf16(row[c * HW + w]);
      for (int64_t c = Cc; c < cs; ++c) out[c] = 0;
    }
  }
}

void decode_sample(mau_tiles* t, Batch* b, int64_t slot, int parts) {
  const int64_t i = b->idx[size_t(slot)];
  const bool flip = !b->flip.empty() && b->flip[size_t(slot)];
  const bool crc = !(t->flags & MAU_TILES_FLAG_NO_CRC);
  Archive mp(t->paths[size_t(i)]);
  Member mem[M_COUNT];
  parse_zip(mp, mem);
  int64_t bytes = 0;
  auto need = [&](int m) -> const Member& {
    if (!mem[m].found) fail(MAU_TILES_E_MEMBER, "'%s': '%.*s' is not a file in the archive", mp.path.c_str(), int(strlen(kMemberNames[m]) - 4), kMemberNames[m]);
    return mem[m];
  };
  const bool zl = (t->flags & MAU_TILES_FLAG_ZLIB) != 0;
  auto image = [&](int m, float* base, const int64_t* d) {
    if (!base) return;
    const Member& M = need(m);
    NpyHeader h = read_member(mp, M, kMemberNames[m], crc, zl, flip ? d[2] : 0, [&](const NpyHeader& hd) {
      if (hd.ndim != 3 || hd.shape[0] != d[0] || hd.shape[1] != d[1] || hd.shape[2] != d[2])
        fail(MAU_TILES_E_SHAPE, "'%s': %s has shape (%lld,%lld,%lld)[ndim %d], the batch expects (%lld,%lld,%lld)", mp.path.c_str(), kMemberNames[m], (long long)hd.shape[0],
             (long long)hd.shape[1], (long long)hd.shape[2], hd.ndim, (long long)d[0], (long long)d[1], (long long)d[2]);
      return base + slot * hd.count;
    });
    bytes += h.count * h.itemsize;
  };
  if (parts & PART_INPUT) {
    image(M_INPUT, b->input, b->dims);
    if (b->input && b->input_staged) {
      const int64_t Cc = b->dims[0], H = b->dims[1], W = b->dims[2];
      stage_tile(b->input + slot * Cc * H * W, Cc, H, W, b->input_staged + slot * H * W * b->staged_cs, b->staged_cs);
    }
  }
  if (parts & PART_REST) {
    image(M_TARGET, b->target, b->dims + 3);
    if (b->metadata) {
      NpyHeader h = read_member(mp, need(M_METADATA), kMemberNames[M_METADATA], crc, zl, 0, [&](const NpyHeader& hd) {
        if (hd.ndim != 1 || hd.shape[0] != b->dims[6])
          fail(MAU_TILES_E_SHAPE, "'%s': metadata has %lld values[ndim %d], the batch expects %lld", mp.path.c_str(), (long long)hd.count, hd.ndim, (long long)b->dims[6]);
        return b->metadata + slot * b->dims[6];
      });
      bytes += h.count * h.itemsize;
    }
    if (b->series) {
      float* dst = b->series + slot * b->series_stride;
      NpyHeader h = read_member(mp, need(M_SERIES), kMemberNames[M_SERIES], crc, zl, 0, [&](const NpyHeader& hd) {
        if (hd.ndim != 1) fail(MAU_TILES_E_SHAPE, "'%s': temperature_serie has %d dimensions, expected 1", mp.path.c_str(), hd.ndim);
        if (hd.count > b->series_stride)
          fail(MAU_TILES_E_CAPACITY, "'%s': temperature_serie has %lld values, buffer holds %lld", mp.path.c_str(), (long long)hd.count, (long long)b->series_stride);
        return dst;
      });
      std::fill(dst + h.count, dst + b->series_stride, 0.f);
      if (b->series_len) b->series_len[slot] = h.count;
      bytes += h.count * h.itemsize;
    }
    t->samples.fetch_add(1, std::memory_order_relaxed);
    t->archive_bytes.fetch_add(int64_t(mp.size), std::memory_order_relaxed);
  }
  t->payload_bytes.fetch_add(bytes, std::memory_order_relaxed);
}

void run_task(mau_tiles* t, std::shared_ptr<Batch> b, int64_t slot, int parts) {
  int code = 0;
  std::string msg;
  bool skip;
  {
    std::lock_guard<std::mutex> g(b->mu);
    skip = b->code != 0;  // the batch already failed: do not spend time on its other samples
  }
  if (!skip) {
    try {
      decode_sample(t, b.get(), slot, parts);
    } catch (const Fail& f) {
      code = f.code;
      msg = f.msg;
    } catch (const std::exception& e) {
      code = MAU_TILES_E_IO;
      msg = e.what();
    }
  }
  std::lock_guard<std::mutex> g(b->mu);
  if (code && !b->code) {
    b->code = code;
    b->msg = msg;
  }
  if (--b->pending == 0) b->cv.notify_all();
}

int set_err(int code, const std::string& m) {
  g_err = m;
  return code;
}

int guard(const std::function<int()>& f) {
  try {
    return f();
  } catch (const Fail& e) {
    return set_err(e.code, e.msg);
  } catch (const std::exception& e) {
    return set_err(MAU_TILES_E_IO, e.what());
  } catch (...) {
    return set_err(MAU_TILES_E_IO, "unknown error");
  }
}

}  // namespace

extern "C" {

const char* mau_tiles_last_error(void) { return g_err.c_str(); }
int mau_tiles_version(void) { return 1; }

int mau_tiles_open(const char* const* paths, int64_t n, int threads, int flags, mau_tiles** out) {
  return guard([&]() -> int {
    if (!out || n < 0 || (n > 0 && !paths)) return set_err(MAU_TILES_E_ARG, "mau_tiles_open: null argument");
    auto t = std::make_unique<mau_tiles>();
    t->paths.reserve(size_t(n));
    for (int64_t i = 0; i < n; ++i) {
      if (!paths[i]) return set_err(MAU_TILES_E_ARG, "mau_tiles_open: null path");
      t->paths.emplace_back(paths[i]);
    }
    t->flags = flags;
    if (threads <= 0) {
      long c = sysconf(_SC_NPROCESSORS_ONLN);
      threads = c > 0 ? int(c) : 4;
    }
    t->pool = std::make_unique<Pool>(std::min(threads, 256), (flags & MAU_TILES_FLAG_NICE) ? 10 : 0);
    *out = t.release();
    return 0;
  });
}

int mau_tiles_close(mau_tiles* t) {
  return guard([&]() -> int {
    if (!t) return 0;
    t->pool.reset();  // drains the queue: buffers of in-flight batches are written before this returns
    delete t;
    return 0;
  });
}

int64_t mau_tiles_count(const mau_tiles* t) { return t ? int64_t(t->paths.size()) : 0; }
int mau_tiles_threads(const mau_tiles* t) { return t && t->pool ? t->pool->size() : 0; }

int mau_tiles_probe(mau_tiles* t, int64_t idx, int64_t dims[8]) {
  return guard([&]() -> int {
    if (!t || !dims) return set_err(MAU_TILES_E_ARG, "mau_tiles_probe: null argument");
    if (idx < 0 || idx >= int64_t(t->paths.size())) return set_err(MAU_TILES_E_ARG, "mau_tiles_probe: index out of range");
    Archive mp(t->paths[size_t(idx)]);
    Member mem[M_COUNT];
    parse_zip(mp, mem);
    const int nd[M_COUNT] = {3, 3, 1, 1};
    const int at[M_COUNT] = {0, 3, 6, 7};
    for (int m = 0; m < M_COUNT; ++m) {
      if (!mem[m].found) fail(MAU_TILES_E_MEMBER, "'%s': '%.*s' is not a file in the archive", mp.path.c_str(), int(strlen(kMemberNames[m]) - 4), kMemberNames[m]);
      MemberStream s(mp, mem[m], kMemberNames[m], false, 16384);  // the header is in the first few hundred bytes
      NpyHeader h = read_npy_header(s, mp, kMemberNames[m]);
      if (h.ndim != nd[m]) fail(MAU_TILES_E_SHAPE, "'%s': %s has %d dimensions, expected %d", mp.path.c_str(), kMemberNames[m], h.ndim, nd[m]);
      for (int k = 0; k < nd[m]; ++k) dims[at[m] + k] = h.shape[k];
    }
    return 0;
  });
}

int mau_tiles_series_lengths(mau_tiles* t, const int64_t* idx, int64_t n, int64_t* out) {
  return guard([&]() -> int {
    if (!t || n < 0 || (n > 0 && (!idx || !out))) return set_err(MAU_TILES_E_ARG, "mau_tiles_series_lengths: null argument");
    for (int64_t k = 0; k < n; ++k)
      if (idx[k] < 0 || idx[k] >= int64_t(t->paths.size())) return set_err(MAU_TILES_E_ARG, "mau_tiles_series_lengths: index " + std::to_string(idx[k]) + " out of range");
    struct Job {
      std::mutex mu;
      std::condition_variable cv;
      int64_t pending;
      int code = 0;
      std::string msg;
    };
    auto job = std::make_shared<Job>();
    job->pending = n;
    for (int64_t k = 0; k < n; ++k) {
      t->pool->push([t, job, k, idx, out] {
        int code = 0;
        std::string msg;
        try {
          Archive mp(t->paths[size_t(idx[k])]);
          Member mem[M_COUNT];
          parse_zip(mp, mem);
          if (!mem[M_SERIES].found) fail(MAU_TILES_E_MEMBER, "'%s': 'temperature_serie' is not a file in the archive", mp.path.c_str());
          MemberStream s(mp, mem[M_SERIES], kMemberNames[M_SERIES], false, 16384);
          NpyHeader h = read_npy_header(s, mp, kMemberNames[M_SERIES]);
          if (h.ndim != 1) fail(MAU_TILES_E_SHAPE, "'%s': temperature_serie has %d dimensions, expected 1", mp.path.c_str(), h.ndim);
          out[k] = h.shape[0];
        } catch (const Fail& f) {
          code = f.code;
          msg = f.msg;
        } catch (const std::exception& e) {
          code = MAU_TILES_E_IO;
          msg = e.what();
        }
        std::lock_guard<std::mutex> g(job->mu);
        if (code && !job->code) {
          job->code = code;
          job->msg = msg;
        }
        if (--job->pending == 0) job->cv.notify_all();
      }, /*urgent=*/true);
    }
    std::unique_lock<std::mutex> g(job->mu);   // idx / out belong to the caller: do not return before every task is done
    job->cv.wait(g, [&] { return job->pending == 0; });
    if (job->code) return set_err(job->code, job->msg);
    return 0;
  });
}

int64_t mau_tiles_submit_staged(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8], float* input,
                                uint16_t* input_staged, int64_t staged_cs, float* target, float* metadata, float* series,
                                int64_t series_stride, int64_t* series_len) {
  int rc = guard([&]() -> int {
    if (!t || !dims || n < 0 || (n > 0 && !idx)) return set_err(MAU_TILES_E_ARG, "mau_tiles_submit: null argument");
    if (series && series_stride <= 0) return set_err(MAU_TILES_E_ARG, "mau_tiles_submit: series_stride must be positive");
    if (input_staged && (!input || staged_cs < dims[0] || staged_cs % 8))
      return set_err(MAU_TILES_E_ARG, "mau_tiles_submit_staged: needs the fp32 input buffer too and a channel stride that is a multiple of 8 and >= C");
    for (int k = 0; k < 7; ++k)
      if (dims[k] < 0) return set_err(MAU_TILES_E_ARG, "mau_tiles_submit: negative dimension");
    for (int64_t k = 0; k < n; ++k)
      if (idx[k] < 0 || idx[k] >= int64_t(t->paths.size())) return set_err(MAU_TILES_E_ARG, "mau_tiles_submit: index " + std::to_string(idx[k]) + " out of range");
    return 0;
  });
  if (rc) return -int64_t(rc);
  auto b = std::make_shared<Batch>();
  b->idx.assign(idx, idx + n);
  if (hflip) b->flip.assign(hflip, hflip + n);
  memcpy(b->dims, dims, sizeof b->dims);
  b->input = input;
  b->input_staged = input_staged;
  b->staged_cs = staged_cs;
  b->target = target;
  b->metadata = metadata;
  b->series = series;
  b->series_stride = series_stride;
  b->series_len = series_len;
  const bool split = input != nullptr && (target || metadata || series);
  b->pending = n * (split ? 2 : 1);
  int64_t ticket;
  {
    std::lock_guard<std::mutex> g(t->mu);
    ticket = t->next_ticket++;
    t->inflight[ticket] = b;
  }
  for (int64_t s = 0; s < n; ++s) {
    if (split) {
      t->pool->push([t, b, s] { run_task(t, b, s, PART_INPUT); });
      t->pool->push([t, b, s] { run_task(t, b, s, PART_REST); });
    } else {
      t->pool->push([t, b, s] { run_task(t, b, s, PART_INPUT | PART_REST); });
    }
  }
  return ticket;
}

int64_t mau_tiles_submit(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8], float* input, float* target,
                         float* metadata, float* series, int64_t series_stride, int64_t* series_len) {
  return mau_tiles_submit_staged(t, idx, n, hflip, dims, input, nullptr, 0, target, metadata, series, series_stride, series_len);
}

int mau_tiles_wait(mau_tiles* t, int64_t ticket) {
  if (!t) return set_err(MAU_TILES_E_ARG, "mau_tiles_wait: null handle");
  std::shared_ptr<Batch> b;
  {
    std::lock_guard<std::mutex> g(t->mu);
    auto it = t->inflight.find(ticket);
    if (it == t->inflight.end()) return set_err(MAU_TILES_E_ARG, "mau_tiles_wait: unknown ticket " + std::to_string(ticket));
    b = it->second;
    t->inflight.erase(it);
  }
  std::unique_lock<std::mutex> g(b->mu);
  b->cv.wait(g, [&] { return b->pending == 0; });
  if (b->code) return set_err(b->code, b->msg);
  return 0;
}

int mau_tiles_done(mau_tiles* t, int64_t ticket) {
  if (!t) return -set_err(MAU_TILES_E_ARG, "mau_tiles_done: null handle");
  std::shared_ptr<Batch> b;
  {
    std::lock_guard<std::mutex> g(t->mu);
    auto it = t->inflight.find(ticket);
    if (it == t->inflight.end()) return -set_err(MAU_TILES_E_ARG, "mau_tiles_done: unknown ticket " + std::to_string(ticket));
    b = it->second;
  }
  std::lock_guard<std::mutex> g(b->mu);
  return b->pending == 0 ? 1 : 0;
}

int mau_tiles_read_batch(mau_tiles* t, const int64_t* idx, int64_t n, const uint8_t* hflip, const int64_t dims[8], float* input, float* target,
                         float* metadata, float* series, int64_t series_stride, int64_t* series_len) {
  int64_t ticket = mau_tiles_submit(t, idx, n, hflip, dims, input, target, metadata, series, series_stride, series_len);
  if (ticket < 0) return int(-ticket);
  return mau_tiles_wait(t, ticket);
}

int mau_tiles_stats(const mau_tiles* t, int64_t* payload_bytes, int64_t* archive_bytes, int64_t* samples) {
  if (!t) return set_err(MAU_TILES_E_ARG, "mau_tiles_stats: null handle");
  if (payload_bytes) *payload_bytes = t->payload_bytes.load();
  if (archive_bytes) *archive_bytes = t->archive_bytes.load();
  if (samples) *samples = t->samples.load();
  return 0;
}

}  // extern "C"

extern "C" int mau_tiles_inflate(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len, size_t split) {
  if (!src || (!dst && dst_len) || split > dst_len) return set_err(MAU_TILES_E_ARG, "mau_tiles_inflate: bad argument");
  try {
    auto inf = std::make_unique<mau_inflate::Inflater>();
    inf->init(src, src_len, src + src_len + 16);
    uint8_t* o = inf->run(dst, dst, dst + split);
    if (o == dst + split) o = inf->run(dst, o, dst + dst_len);
    if (o != dst + dst_len) return set_err(MAU_TILES_E_FORMAT, "inflate: stream ends early");
    if (!inf->done()) return set_err(MAU_TILES_E_FORMAT, "inflate: more data than dst_len");
    return 0;
  } catch (const mau_inflate::Error& e) {
    return set_err(MAU_TILES_E_FORMAT, e.what);
  }
}

extern "C" uint32_t mau_tiles_crc32(uint32_t crc, const void* data, size_t n) { return crc32_update(crc, static_cast<const uint8_t*>(data), n); }

// ---- repack: the same archive with stored (uncompressed) members ---------------------------------------------
namespace {

void put16(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back(uint8_t(x));
  v.push_back(uint8_t(x >> 8));
}
void put32(std::vector<uint8_t>& v, uint32_t x) {
  put16(v, x & 0xFFFFu);
  put16(v, x >> 16);
}

// all bytes of one member, CRC verified
void member_bytes(const Archive& mp, const Member& M, const std::string& name, std::vector<uint8_t>& out) {
  out.resize(size_t(M.usize));
  if (M.method == 0) {
    if (M.csize != M.usize) fail(MAU_TILES_E_FORMAT, "'%s': stored member %s has differing sizes", mp.path.c_str(), name.c_str());
    if (M.usize) mp.read_at(out.data(), M.data_off, M.usize);
  } else {
    try {
      const uint8_t* in = mp.fetch(M.data_off, M.csize, 16);
      auto inf = std::make_unique<mau_inflate::Inflater>();
      inf->init(in, size_t(M.csize), in + M.csize + 16);
      uint8_t* end = out.data() + out.size();
      if (inf->run(out.data(), out.data(), end) != end || !inf->done())
        fail(MAU_TILES_E_FORMAT, "'%s': member %s does not decode to the size its directory entry gives", mp.path.c_str(), name.c_str());
    } catch (const mau_inflate::Error& e) {
      fail(MAU_TILES_E_FORMAT, "'%s': error while decompressing %s (%s)", mp.path.c_str(), name.c_str(), e.what);
    }
  }
  if (crc32_update(0, out.data(), out.size()) != M.crc) fail(MAU_TILES_E_FORMAT, "'%s': bad CRC-32 for member %s", mp.path.c_str(), name.c_str());
}

void write_all(int fd, const void* p, size_t n, const std::string& path) {
  const uint8_t* q = static_cast<const uint8_t*>(p);
  while (n) {
    ssize_t w = ::write(fd, q, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      fail(MAU_TILES_E_IO, "cannot write '%s': %s", path.c_str(), strerror(errno));
    }
    q += w;
    n -= size_t(w);
  }
}

}  // namespace

extern "C" int mau_tiles_repack(const char* src_path, const char* dst_path) {
  return guard([&]() -> int {
    if (!src_path || !dst_path) return set_err(MAU_TILES_E_ARG, "mau_tiles_repack: null path");
    Archive mp(src_path);
    struct Item {
      std::string name;
      Member m;
    };
    std::vector<Item> items;
    walk_zip(mp, [&](const char* name, size_t nlen, const Member& M) {
      check_supported(mp, M, name, nlen);
      if (M.usize >= 0xFFFFFFFFull) fail(MAU_TILES_E_FORMAT, "'%s': member %.*s is too large for a plain ZIP entry", mp.path.c_str(), int(nlen), name);
      items.push_back(Item{std::string(name, nlen), M});
    });
    const std::string dst(dst_path), tmp = dst + ".tmp" + std::to_string(long(getpid())) + "." + std::to_string(long(syscall(SYS_gettid)));
    int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd < 0) fail(MAU_TILES_E_IO, "cannot create '%s': %s", tmp.c_str(), strerror(errno));
    try {
      std::vector<uint8_t> central, head, body;
      uint64_t offset = 0;
      for (const Item& it : items) {
        member_bytes(mp, it.m, it.name, body);
        if (offset + 30 + it.name.size() + body.size() >= 0xFFFFFFFFull) fail(MAU_TILES_E_FORMAT, "'%s': repacked archive would need ZIP64", mp.path.c_str());
        head.clear();
        put32(head, 0x04034b50u);
        put16(head, 20);               // version needed
        put16(head, 0);                // flags
        put16(head, 0);                // method: stored
        put32(head, it.m.dos_time_date);
        put32(head, it.m.crc);
        put32(head, uint32_t(body.size()));
        put32(head, uint32_t(body.size()));
        put16(head, uint32_t(it.name.size()));
        put16(head, 0);                // no extra field
        head.insert(head.end(), it.name.begin(), it.name.end());
        write_all(fd, head.data(), head.size(), tmp);
        write_all(fd, body.data(), body.size(), tmp);
        put32(central, 0x02014b50u);
        put16(central, 20);            // version made by
        put16(central, 20);
        put16(central, 0);
        put16(central, 0);
        put32(central, it.m.dos_time_date);
        put32(central, it.m.crc);
        put32(central, uint32_t(body.size()));
        put32(central, uint32_t(body.size()));
        put16(central, uint32_t(it.name.size()));
        put16(central, 0);             // extra
        put16(central, 0);             // comment
        put16(central, 0);             // disk
        put16(central, 0);             // internal attributes
        put32(central, 0);             // external attributes
        put32(central, uint32_t(offset));
        central.insert(central.end(), it.name.begin(), it.name.end());
        offset += head.size() + body.size();
      }
      if (items.size() >= 0xFFFFu) fail(MAU_TILES_E_FORMAT, "'%s': too many members", mp.path.c_str());
      std::vector<uint8_t> eocd;
      put32(eocd, 0x06054b50u);
      put16(eocd, 0);
      put16(eocd, 0);
      put16(eocd, uint32_t(items.size()));
      put16(eocd, uint32_t(items.size()));
      put32(eocd, uint32_t(central.size()));
      put32(eocd, uint32_t(offset));
      put16(eocd, 0);
      write_all(fd, central.data(), central.size(), tmp);
      write_all(fd, eocd.data(), eocd.size(), tmp);
      if (::close(fd) != 0) {
        fd = -1;
        fail(MAU_TILES_E_IO, "cannot close '%s': %s", tmp.c_str(), strerror(errno));
      }
      fd = -1;
      if (::rename(tmp.c_str(), dst.c_str()) != 0) fail(MAU_TILES_E_IO, "cannot rename '%s' to '%s': %s", tmp.c_str(), dst.c_str(), strerror(errno));
    } catch (...) {
      if (fd >= 0) ::close(fd);
      ::unlink(tmp.c_str());
      throw;
    }
    return 0;
  });
}
