// Alphabet profiles shared by host and device code.
//
// Restates the *behaviour* of the reference's profiles (table contents are
// fixed by the IUPAC standard and by the reference's observable matching
// rules, cited per item) in the form the CUDA path needs: a text byte is
// reduced to a small "row" index, and every query carries one equality
// bit-vector per row.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif

namespace sb {

enum Profile : int { kDna = 0, kIupac = 1, kAscii = 2 };

// Number of equality rows a text byte can select, and how a byte maps to one.
//  Dna:   row = (c >> 1) & 3        (reference src/profiles/dna.rs:19-23,135-137;
//                                    every byte is accepted, case-insensitive)
//  Iupac: row = c & 0x1F            (reference src/profiles/iupac.rs:146-148; the
//                                    5 low ASCII bits, so case-insensitive)
template <int P> struct ProfileTraits;
template <> struct ProfileTraits<kDna> {
  static constexpr int kRows = 4;
  static constexpr int kShift = 1;
  static constexpr uint32_t kMask = 3;
};
template <> struct ProfileTraits<kIupac> {
  static constexpr int kRows = 32;
  static constexpr int kShift = 0;
  static constexpr uint32_t kMask = 31;
};
//  Ascii: row = c                   (reference src/profiles/ascii.rs:13-73, the default
//                                    case-sensitive Ascii<true> the C ABI uses, src/c.rs:63)
template <> struct ProfileTraits<kAscii> {
  static constexpr int kRows = 256;
  static constexpr int kShift = 0;
  static constexpr uint32_t kMask = 255;
};

// 4-bit set of {A=1,C=2,T=4,G=8} for the letter whose low 5 ASCII bits are i;
// 0 for X (matches nothing), 255 for bytes that are not IUPAC letters
// (reference src/profiles/iupac.rs:281-317).
SB_HD uint8_t iupac_code(uint8_t c) {
  switch (c & 31) {
    case 'A' & 31: return 1;
    case 'C' & 31: return 2;
    case 'T' & 31: return 4;
    case 'U' & 31: return 4;
    case 'G' & 31: return 8;
    case 'N' & 31: return 15;
    case 'R' & 31: return 1 | 8;
    case 'Y' & 31: return 2 | 4;
    case 'S' & 31: return 8 | 2;
    case 'W' & 31: return 1 | 4;
    case 'K' & 31: return 8 | 4;
    case 'M' & 31: return 1 | 2;
    case 'B' & 31: return 2 | 8 | 4;
    case 'D' & 31: return 1 | 8 | 4;
    case 'H' & 31: return 1 | 2 | 4;
    case 'V' & 31: return 1 | 2 | 8;
    case 'X' & 31: return 0;
    default: return 255;
  }
}

// Does pattern byte p match a text byte that selects `row`, as seen by the
// *search* DP?  Iupac text bytes outside the table behave as N: only the low
// nibble of their code (255 -> 0xF) takes part (src/profiles/iupac.rs:68-128).
template <int P> SB_HD bool row_matches(uint8_t p, int row);
template <> SB_HD bool row_matches<kDna>(uint8_t p, int row) { return ((p >> 1) & 3) == row; }
template <> SB_HD bool row_matches<kIupac>(uint8_t p, int row) {
  return (iupac_code(p) & (iupac_code((uint8_t)row) & 0x0F)) != 0;
}

template <> SB_HD bool row_matches<kAscii>(uint8_t p, int row) { return (int)p == row; }

// Profile::is_match, used by the traceback only.
//  Dna:   (a|0x20) == (b|0x20)          src/profiles/dna.rs:48-50
//  Iupac: code(a) & code(b) != 0        src/profiles/iupac.rs:136-138
//  Ascii: a == b                        src/profiles/ascii.rs:44-51
template <int P> SB_HD bool trace_match(uint8_t p, uint8_t t);
template <> SB_HD bool trace_match<kDna>(uint8_t p, uint8_t t) { return (p | 0x20) == (t | 0x20); }
template <> SB_HD bool trace_match<kIupac>(uint8_t p, uint8_t t) {
  return (iupac_code(p) & iupac_code(t)) != 0;
}

template <> SB_HD bool trace_match<kAscii>(uint8_t p, uint8_t t) { return p == t; }  // ascii.rs:44-51

// Iupac::valid_seq (scalar branch, src/profiles/iupac.rs:195-201).
inline bool iupac_valid(const uint8_t* s, size_t n) {
  for (size_t i = 0; i < n; i++) {
    uint8_t c = s[i] & (uint8_t)~0x20;
    if (c <= '@' || c >= 'Z' || iupac_code(c) == 255) return false;
  }
  return true;
}

// Complement of one byte.  Dna only maps upper-case ACGT (src/profiles/dna.rs:
// 121-133); Iupac maps the 16 letters in both cases (src/profiles/iupac.rs:235-278);
// every other byte is returned unchanged.
inline uint8_t complement_byte(int profile, uint8_t c) {
  if (profile == kDna) {
    switch (c) {
      case 'A': return 'T';
      case 'C': return 'G';
      case 'T': return 'A';
      case 'G': return 'C';
      default: return c;
    }
  }
  const char* from = "ACTGRYSWKMBDHVNX";
  const char* to = "TGACYRSWMKVHDBNX";
  uint8_t up = c & (uint8_t)~0x20;
  for (int i = 0; from[i]; i++)
    if (up == (uint8_t)from[i] && ((c | 0x20) >= 'a' && (c | 0x20) <= 'z'))
      return (uint8_t)(to[i] | (c & 0x20));
  return c;
}

}  // namespace sb
