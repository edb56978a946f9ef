// ORACLE (test infrastructure) — LSB-first bit writer / reader (JPEG XL packs bits
// least-significant first inside each byte; ISO/IEC 18181-1 section "bitstream") and the
// field codings used by the headers (U32 selectors, U64, VarLenUint8/16) [UPSTREAM].
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace jxo {

class BitWriter {
 public:
  void Write(int nbits, uint64_t value) {   // nbits <= 56
    for (int i = 0; i < nbits; ++i) {
      const size_t byte = pos_ >> 3;
      if (byte >= data_.size()) data_.push_back(0);
      if ((value >> i) & 1) data_[byte] |= (uint8_t)(1u << (pos_ & 7));
      ++pos_;
    }
  }
  void ZeroPadToByte() { pos_ = (pos_ + 7) & ~(size_t)7; while (data_.size() < (pos_ >> 3)) data_.push_back(0); }
  size_t BitsWritten() const { return pos_; }
  const std::vector<uint8_t>& Bytes() { while (data_.size() < ((pos_ + 7) >> 3)) data_.push_back(0); return data_; }
  void AppendBytes(const std::vector<uint8_t>& b) {  // requires byte alignment
    ZeroPadToByte();
    data_.insert(data_.end(), b.begin(), b.end());
    pos_ += b.size() * 8;
  }
  // JPEG XL U32 with the four-way selector written explicitly by the caller
  void WriteSelector(int sel) { Write(2, (uint64_t)sel); }
  void WriteU64(uint64_t v) {
    if (v == 0) { Write(2, 0); }
    else if (v <= 16) { Write(2, 1); Write(4, v - 1); }
    else if (v <= 272) { Write(2, 2); Write(8, v - 17); }
    else {
      Write(2, 3); Write(12, v & 4095); v >>= 12;
      int shift = 12;
      while (v > 0 && shift < 60) { Write(1, 1); Write(8, v & 255); v >>= 8; shift += 8; }
      if (shift < 60) Write(1, 0); else { /* last 4 bits */ Write(4, v & 15); }
    }
  }
  void WriteVarLenUint8(uint32_t n) {   // n in [0, 255]
    if (n == 0) { Write(1, 0); return; }
    Write(1, 1);
    int nbits = 31 - __builtin_clz(n);
    Write(3, (uint64_t)nbits);
    Write(nbits, n - (1u << nbits));
  }
  void WriteVarLenUint16(uint32_t n) {  // n in [0, 65535]
    if (n == 0) { Write(1, 0); return; }
    Write(1, 1);
    int nbits = 31 - __builtin_clz(n);
    Write(4, (uint64_t)nbits);
    Write(nbits, n - (1u << nbits));
  }

 private:
  std::vector<uint8_t> data_;
  size_t pos_ = 0;
};

class BitReader {
 public:
  BitReader(const uint8_t* d, size_t n) : d_(d), n_(n) {}
  uint64_t Read(int nbits) {
    uint64_t v = 0;
    for (int i = 0; i < nbits; ++i) {
      const size_t byte = pos_ >> 3;
      uint64_t bit = 0;
      if (byte < n_) bit = (d_[byte] >> (pos_ & 7)) & 1; else overrun_ = true;
      v |= bit << i;
      ++pos_;
    }
    return v;
  }
  uint64_t Peek(int nbits) { const size_t p = pos_; const bool o = overrun_; uint64_t v = Read(nbits); pos_ = p; overrun_ = o; return v; }
  void Skip(int nbits) { pos_ += nbits; }
  void ZeroPadToByte() { pos_ = (pos_ + 7) & ~(size_t)7; }
  size_t Pos() const { return pos_; }
  void Seek(size_t bitpos) { pos_ = bitpos; }
  bool Overrun() const { return overrun_ || pos_ > n_ * 8; }
  uint64_t ReadU64() {
    const int sel = (int)Read(2);
    if (sel == 0) return 0;
    if (sel == 1) return 1 + Read(4);
    if (sel == 2) return 17 + Read(8);
    uint64_t v = Read(12);
    int shift = 12;
    while (Read(1)) {
      if (shift == 60) { v |= Read(4) << shift; break; }
      v |= Read(8) << shift;
      shift += 8;
    }
    return v;
  }
  uint32_t ReadVarLenUint8() {
    if (!Read(1)) return 0;
    const int nbits = (int)Read(3);
    if (nbits == 0) return 1;
    return (uint32_t)Read(nbits) + (1u << nbits);
  }
  uint32_t ReadVarLenUint16() {
    if (!Read(1)) return 0;
    const int nbits = (int)Read(4);
    if (nbits == 0) return 1;
    return (uint32_t)Read(nbits) + (1u << nbits);
  }

 private:
  const uint8_t* d_;
  size_t n_;
  size_t pos_ = 0;
  bool overrun_ = false;
};

}  // namespace jxo
