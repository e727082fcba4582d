// inflate_fast.h -- raw DEFLATE (RFC 1951) decoder for the tile reader.
//
// The reference's archives are np.savez_compressed members (zlib level 6, reference
// src/data/processing_10m/process.py:187); np.load inflates them through zlib one 6 MB member at a time on
// the training thread.  Decoding is the loader's bottleneck (80 % of a sample's time with zlib's inflate), so
// the reader carries its own decoder built for one shot, whole-member decoding straight into the destination:
//   * 64-bit bit buffer with a branch-free refill (one unaligned 8-byte load per refill), up to three literals
//     per refill;
//   * two-level canonical-Huffman tables (11-bit primary for literal/length, 8-bit for distance) with the
//     base value and the number of extra bits folded into the entry;
//   * the output buffer itself is the window -- no sliding-window copy; matches are copied with 8-byte words,
//     short distances (runs of zeros / repeated fp32 words of the one-hot Dynamic World planes) by storing the
//     period replicated over a 64-bit word;
//   * resumable at an output boundary (pending match kept in the state), so the NPY header can be decoded into
//     a scratch buffer and the payload into its slot of the batch.
// Measured and dropped: table entries that deliver two literals when both codes fit the primary index (1.6-1.7x on
// streams of <= 5-bit literal codes, 1.2x on Huffman-only text, but 0.97x on the tiles: the fp32 planes' 8-9 bit
// codes never pair and every literal pays for the wider store).
// Output is bit-identical to zlib's (tests/test_tiles_cpu.py compares against zlib on every block type).
#ifndef MAU_INFLATE_FAST_H_
#define MAU_INFLATE_FAST_H_

#include <cstddef>
#include <cstdint>
#include <cstring>

namespace mau_inflate {

struct Error {
  const char* what;
};

// Table entry, one 32-bit word so that the hot loop tests single bits:
//   bits  0..7   code bits to consume at this level
//   bits  8..12  number of extra bits (K_BASE) or index bits of the sub-table (K_SUB)
//   bits 13..14  kind of a non-literal entry, bit 15 = "exceptional" (sub-table pointer, end of block, hole)
//   bits 16..30  literal / base length / base distance / sub-table offset
//   bit  31      literal
typedef uint32_t Entry;
constexpr Entry E_LITERAL = 1u << 31, E_EXC = 1u << 15;
constexpr Entry E_SUB = E_EXC | (0u << 13), E_EOB = E_EXC | (1u << 13), E_INVALID = E_EXC | (2u << 13);
constexpr Entry E_EXC_KIND = E_EXC | (3u << 13);
inline Entry mk(Entry flags, uint32_t val, uint32_t extra, uint32_t len) { return flags | (val << 16) | (extra << 8) | len; }
inline uint32_t e_len(Entry e) { return e & 0xFF; }
inline uint32_t e_extra(Entry e) { return (e >> 8) & 31; }
inline uint32_t e_val(Entry e) { return (e >> 16) & 0x7FFF; }

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitTableSize = (1 << kLitBits) + 288 * 16;    // 15-bit codes: sub-tables of <= 2^4 entries
constexpr int kDistTableSize = (1 << kDistBits) + 32 * 128;  // sub-tables of <= 2^7 entries

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// Builds a two-level decode table from code lengths.  `make(sym)` gives the entry payload for a symbol.
// Returns false for an over-subscribed code; incomplete codes leave K_INVALID holes (an error only if hit).
template <typename Make>
bool build_table(const uint8_t* lens, int nsym, int primary_bits, Entry* table, int table_cap, Make make) {
  int count[16] = {0};
  for (int s = 0; s < nsym; ++s) count[lens[s]]++;
  count[0] = 0;
  int left = 1;
  for (int l = 1; l <= 15; ++l) {
    left = (left << 1) - count[l];
    if (left < 0) return false;  // over-subscribed
  }
  uint32_t next_code[16];
  uint32_t code = 0;
  for (int l = 1; l <= 15; ++l) {
    code = (code + uint32_t(count[l - 1])) << 1;
    next_code[l] = code;
  }
  const int psize = 1 << primary_bits;
  for (int i = 0; i < psize; ++i) table[i] = E_INVALID;
  // pass 1: longest code behind every primary prefix that needs a sub-table
  uint8_t sub_bits[1 << kLitBits] = {0};
  uint32_t codes[288 + 32];
  for (int s = 0; s < nsym; ++s) {
    int l = lens[s];
    if (!l) continue;
    uint32_t rev = reverse_bits(next_code[l]++, l);
    codes[s] = rev;
    if (l > primary_bits) {
      uint32_t p = rev & uint32_t(psize - 1);
      if (l - primary_bits > sub_bits[p]) sub_bits[p] = uint8_t(l - primary_bits);
    }
  }
  int used = psize;
  for (int p = 0; p < psize; ++p) {
    if (!sub_bits[p]) continue;
    int n = 1 << sub_bits[p];
    if (used + n > table_cap) return false;
    table[p] = mk(E_SUB, uint32_t(used), sub_bits[p], uint32_t(primary_bits));
    for (int i = 0; i < n; ++i) table[used + i] = E_INVALID;
    used += n;
  }
  // pass 2: fill
  for (int s = 0; s < nsym; ++s) {
    int l = lens[s];
    if (!l) continue;
    uint32_t rev = codes[s];
    Entry e = make(s);
    if (l <= primary_bits) {
      e |= uint32_t(l);
      for (uint32_t i = rev; i < uint32_t(psize); i += 1u << l) table[i] = e;
    } else {
      uint32_t p = rev & uint32_t(psize - 1);
      int sb = sub_bits[p], rl = l - primary_bits;
      Entry* sub = table + e_val(table[p]);
      e |= uint32_t(rl);
      for (uint32_t i = rev >> primary_bits; i < (1u << sb); i += 1u << rl) sub[i] = e;
    }
  }
  return true;
}

inline Entry make_litlen(int s) {
  static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  if (s < 256) return mk(E_LITERAL, uint32_t(s), 0, 0);
  if (s == 256) return mk(E_EOB, 0, 0, 0);
  if (s > 285) return E_INVALID;
  return mk(0, base[s - 257], extra[s - 257], 0);
}

inline Entry make_dist(int s) {
  static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  if (s > 29) return E_INVALID;
  return mk(0, base[s], extra[s], 0);
}

class Inflater {
 public:
  // [in, in + n) is the raw deflate stream; at least 16 readable bytes must follow it (`readable_end`).
  void init(const uint8_t* in, size_t n, const uint8_t* readable_end) {
    in_begin_ = in_next_ = in;
    in_end_ = in + n;
    if (size_t(readable_end - in_end_) < 16) throw Error{"inflate: input needs 16 readable bytes of slack"};
    in_stop_ = in_end_ + 8;  // an 8-byte load at in_stop_ still ends inside the slack
    bitbuf_ = 0;
    bitcnt_ = 0;
    state_ = S_HEADER;
    final_ = false;
    match_left_ = 0;
  }
  bool done() const { return state_ == S_DONE; }
  size_t consumed() const { return size_t(in_next_ - in_begin_) - (bitcnt_ >> 3); }

  // Decodes into [out, out_end); [out_begin, out) is the history matches may reach into.  Returns the new
  // output position: out_end (more output pending, call again) or the end of the stream (done()).
  uint8_t* run(uint8_t* out_begin, uint8_t* out, uint8_t* out_end) {
    if (match_left_) {
      out = copy_careful(out_begin, out, out_end);
      if (match_left_) return out;
    }
    for (;;) {
      switch (state_) {
        case S_DONE:
          if (consumed() > size_t(in_end_ - in_begin_)) throw Error{"inflate: stream runs past the end of the member"};
          return out;
        case S_HEADER: {
          refill();
          final_ = bitbuf_ & 1;
          uint32_t type = (bitbuf_ >> 1) & 3;
          drop(3);
          if (type == 0) {
            drop(bitcnt_ & 7);  // to the byte boundary; the bit buffer now holds whole bytes
            in_next_ -= bitcnt_ >> 3;
            bitbuf_ = 0;
            bitcnt_ = 0;
            if (in_end_ - in_next_ < 4) throw Error{"inflate: truncated stored block"};
            uint32_t len = in_next_[0] | (in_next_[1] << 8), nlen = in_next_[2] | (in_next_[3] << 8);
            if ((len ^ nlen) != 0xFFFFu) throw Error{"inflate: stored block length check failed"};
            in_next_ += 4;
            stored_left_ = len;
            state_ = S_STORED;
          } else if (type == 1) {
            fixed_tables();
            state_ = S_HUFF;
          } else if (type == 2) {
            dynamic_tables();
            state_ = S_HUFF;
          } else {
            throw Error{"inflate: invalid block type"};
          }
          break;
        }
        case S_STORED: {
          size_t n = stored_left_;
          if (n > size_t(out_end - out)) n = size_t(out_end - out);
          if (n > size_t(in_end_ - in_next_)) throw Error{"inflate: truncated stored block"};
          memcpy(out, in_next_, n);
          out += n;
          in_next_ += n;
          stored_left_ -= uint32_t(n);
          if (stored_left_) return out;  // output full
          state_ = final_ ? S_DONE : S_HEADER;
          break;
        }
        case S_HUFF: {
          out = huff_fast(out_begin, out, out_end);
          bool full = false;
          out = huff_careful(out_begin, out, out_end, &full);
          if (full) return out;
          break;
        }
      }
    }
  }

 private:
  enum State { S_HEADER, S_STORED, S_HUFF, S_DONE };

  static uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64 / aarch64)
  }
  void refill() {
    if (in_next_ > in_stop_) throw Error{"inflate: stream runs past the end of the member"};
    bitbuf_ |= load64(in_next_) << bitcnt_;
    in_next_ += (63 - bitcnt_) >> 3;
    bitcnt_ |= 56;
  }
  void drop(uint32_t n) {
    bitbuf_ >>= n;
    bitcnt_ -= n;
  }
  uint32_t take(uint32_t n) {
    uint32_t v = uint32_t(bitbuf_ & ((uint64_t(1) << n) - 1));
    drop(n);
    return v;
  }

  void fixed_tables() {
    uint8_t l[288], d[32];
    for (int i = 0; i < 144; ++i) l[i] = 8;
    for (int i = 144; i < 256; ++i) l[i] = 9;
    for (int i = 256; i < 280; ++i) l[i] = 7;
    for (int i = 280; i < 288; ++i) l[i] = 8;
    for (int i = 0; i < 32; ++i) d[i] = 5;
    build_table(l, 288, kLitBits, lit_, kLitTableSize, make_litlen);
    build_table(d, 32, kDistBits, dist_, kDistTableSize, make_dist);
  }

  void dynamic_tables() {
    refill();
    uint32_t hlit = take(5) + 257, hdist = take(5) + 1, hclen = take(4) + 4;
    if (hlit > 286 || hdist > 30) throw Error{"inflate: too many length or distance symbols"};
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t pl[19] = {0};
    for (uint32_t i = 0; i < hclen; ++i) {
      if (bitcnt_ < 3) refill();
      pl[order[i]] = uint8_t(take(3));
    }
    Entry pre[1 << kPreBits];
    if (!build_table(pl, 19, kPreBits, pre, 1 << kPreBits, [](int s) { return mk(E_LITERAL, uint32_t(s), 0, 0); }))
      throw Error{"inflate: invalid code lengths set"};
    uint8_t lens[288 + 32];
    uint32_t n = 0, total = hlit + hdist;
    while (n < total) {
      refill();
      Entry e = pre[bitbuf_ & ((1u << kPreBits) - 1)];
      if (!(e & E_LITERAL)) throw Error{"inflate: invalid code lengths set"};
      drop(e_len(e));
      uint32_t sym = e_val(e);
      if (sym < 16) {
        lens[n++] = uint8_t(sym);
        continue;
      }
      uint32_t rep;
      uint8_t v = 0;
      if (sym == 16) {
        if (n == 0) throw Error{"inflate: invalid bit length repeat"};
        v = lens[n - 1];
        rep = 3 + take(2);
      } else if (sym == 17) {
        rep = 3 + take(3);
      } else {
        rep = 11 + take(7);
      }
      if (n + rep > total) throw Error{"inflate: invalid bit length repeat"};
      while (rep--) lens[n++] = v;
    }
    if (lens[256] == 0) throw Error{"inflate: invalid code -- missing end-of-block"};
    if (!build_table(lens, int(hlit), kLitBits, lit_, kLitTableSize, make_litlen)) throw Error{"inflate: invalid literal/lengths set"};
    if (!build_table(lens + hlit, int(hdist), kDistBits, dist_, kDistTableSize, make_dist)) throw Error{"inflate: invalid distances set"};
  }

  // decode loop with no bounds checks inside: needs kFastRoom bytes of output room and input short of the slack
  static constexpr ptrdiff_t kFastRoom = 8 + 258 + 40;
  uint8_t* huff_fast(uint8_t* out_begin, uint8_t* out, uint8_t* out_end) {
    const uint64_t lmask = (1u << kLitBits) - 1, dmask = (1u << kDistBits) - 1;
    uint64_t bb = bitbuf_;
    uint32_t bc = bitcnt_;
    const uint8_t* in = in_next_;
    const Entry* const lit = lit_;
    const Entry* const dst = dist_;
#define MAU_REFILL()                 \
  do {                               \
    bb |= load64(in) << bc;          \
    in += (63 - bc) >> 3;            \
    bc |= 56;                        \
  } while (0)
#define MAU_DROP(n)  \
  do {               \
    bb >>= (n);      \
    bc -= (n);       \
  } while (0)
    while (out_end - out >= kFastRoom && in <= in_end_) {
      MAU_REFILL();  // >= 56 bits: 3 x 11 (literals) + 15 + 5 (last code and its extra bits) fit
      Entry e = lit[bb & lmask];
      if (e & E_LITERAL) {  // up to three primary-table literals (<= 11 bits each) ahead of the code handled below
        MAU_DROP(e & 0xFF);
        *out++ = uint8_t(e >> 16);
        e = lit[bb & lmask];
        if (e & E_LITERAL) {
          MAU_DROP(e & 0xFF);
          *out++ = uint8_t(e >> 16);
          e = lit[bb & lmask];
          if (e & E_LITERAL) {
            MAU_DROP(e & 0xFF);
            *out++ = uint8_t(e >> 16);
            e = lit[bb & lmask];
          }
        }
      }
      if (e & E_EXC) {  // rare: a code longer than the primary table, end of block, or a hole
        if ((e & E_EXC_KIND) == E_SUB) {
          MAU_DROP(e & 0xFF);
          e = lit[e_val(e) + (bb & ((1u << e_extra(e)) - 1))];
        }
        if ((e & E_EXC_KIND) == E_EOB) {
          MAU_DROP(e & 0xFF);
          bitbuf_ = bb;
          bitcnt_ = bc;
          in_next_ = in;
          state_ = final_ ? S_DONE : S_HEADER;
          return out;
        }
        if (e & E_EXC) throw Error{"inflate: invalid literal/length code"};
      }
      MAU_DROP(e & 0xFF);
      if (e & E_LITERAL) {
        *out++ = uint8_t(e >> 16);
        continue;
      }
      uint32_t xb = e_extra(e);
      uint32_t len = e_val(e) + uint32_t(bb & ((1u << xb) - 1));
      MAU_DROP(xb);
      MAU_REFILL();  // distance code <= 15 bits + 13 extra bits
      e = dst[bb & dmask];
      if (e & E_EXC) {
        if ((e & E_EXC_KIND) != E_SUB) throw Error{"inflate: invalid distance code"};
        MAU_DROP(e & 0xFF);
        e = dst[e_val(e) + (bb & ((1u << e_extra(e)) - 1))];
        if (e & E_EXC) throw Error{"inflate: invalid distance code"};
      }
      MAU_DROP(e & 0xFF);
      xb = e_extra(e);
      uint32_t dist = e_val(e) + uint32_t(bb & ((1u << xb) - 1));
      MAU_DROP(xb);
      if (dist > size_t(out - out_begin)) throw Error{"inflate: invalid distance too far back"};
      // match copy with 8-byte words; may write up to 40 bytes past out + len (room guaranteed by the loop condition)
      uint8_t* o = out;
      uint8_t* const oe = out + len;
      const uint8_t* s = out - dist;
      if (dist >= 8) {
        memcpy(o, s, 8);
        memcpy(o + 8, s + 8, 8);
        memcpy(o + 16, s + 16, 8);
        memcpy(o + 24, s + 24, 8);
        if (len > 32) {
          o += 32;
          s += 32;
          do {
            memcpy(o, s, 8);
            memcpy(o + 8, s + 8, 8);
            o += 16;
            s += 16;
          } while (o < oe);
        }
      } else {
        // period < 8 (runs of one byte, repeated fp32 words of the one-hot planes): replicate the period over one
        // 64-bit word and store it at a stride that is a multiple of the period
        static const uint8_t stride[8] = {0, 8, 8, 6, 8, 5, 6, 7};
        uint64_t v = load64(s);  // the first `dist` bytes are history, the rest is overwritten below
        uint32_t sh = dist * 8;
        v &= (uint64_t(1) << sh) - 1;
        v |= v << sh;
        if (sh < 32) {
          v |= v << (2 * sh);
          if (sh < 16) v |= v << (4 * sh);
        }
        const uint32_t st = stride[dist];
        do {
          memcpy(o, &v, 8);
          o += st;
        } while (o < oe);
      }
      out = oe;
    }
#undef MAU_REFILL
#undef MAU_DROP
    bitbuf_ = bb;
    bitcnt_ = bc;
    in_next_ = in;
    return out;
  }

  // one code: sub-table resolved, bits of the first level dropped, e_len(e) bits still to drop
  Entry lookup(const Entry* T, uint32_t primary_bits) {
    Entry e = T[bitbuf_ & ((1u << primary_bits) - 1)];
    if ((e & E_EXC_KIND) == E_SUB) {
      drop(e_len(e));
      e = T[e_val(e) + (bitbuf_ & ((1u << e_extra(e)) - 1))];
    }
    return e;
  }

  // symbol-at-a-time loop with exact bounds; stops (full = true) when the next symbol would produce output
  // beyond out_end, leaving that symbol unconsumed (or a partially copied match pending)
  uint8_t* huff_careful(uint8_t* out_begin, uint8_t* out, uint8_t* out_end, bool* full) {
    while (state_ == S_HUFF) {
      refill();
      uint64_t save_bb = bitbuf_;
      uint32_t save_bc = bitcnt_;
      Entry e = lookup(lit_, kLitBits);
      if ((e & E_EXC_KIND) == E_EOB) {
        drop(e_len(e));
        state_ = final_ ? S_DONE : S_HEADER;
        return out;
      }
      if (e & E_EXC) throw Error{"inflate: invalid literal/length code"};
      if (out == out_end) {
        bitbuf_ = save_bb;
        bitcnt_ = save_bc;
        *full = true;
        return out;
      }
      drop(e_len(e));
      if (e & E_LITERAL) {
        *out++ = uint8_t(e >> 16);
        continue;
      }
      uint32_t len = e_val(e) + take(e_extra(e));
      refill();
      e = lookup(dist_, kDistBits);
      if (e & (E_EXC | E_LITERAL)) throw Error{"inflate: invalid distance code"};
      drop(e_len(e));
      uint32_t dist = e_val(e) + take(e_extra(e));
      if (dist > size_t(out - out_begin)) throw Error{"inflate: invalid distance too far back"};
      match_left_ = len;
      match_dist_ = dist;
      out = copy_careful(out_begin, out, out_end);
      if (match_left_) {
        *full = true;
        return out;
      }
    }
    return out;
  }

  uint8_t* copy_careful(uint8_t* out_begin, uint8_t* out, uint8_t* out_end) {
    if (match_dist_ > size_t(out - out_begin)) throw Error{"inflate: invalid distance too far back"};
    size_t n = match_left_;
    if (n > size_t(out_end - out)) n = size_t(out_end - out);
    const uint8_t* s = out - match_dist_;
    for (size_t i = 0; i < n; ++i) out[i] = s[i];
    match_left_ -= uint32_t(n);
    return out + n;
  }

  const uint8_t *in_begin_ = nullptr, *in_next_ = nullptr, *in_end_ = nullptr, *in_stop_ = nullptr;
  uint64_t bitbuf_ = 0;
  uint32_t bitcnt_ = 0;
  State state_ = S_HEADER;
  bool final_ = false;
  uint32_t stored_left_ = 0, match_left_ = 0, match_dist_ = 0;
  Entry lit_[kLitTableSize];
  Entry dist_[kDistTableSize];
};

}  // namespace mau_inflate
#endif  // MAU_INFLATE_FAST_H_
