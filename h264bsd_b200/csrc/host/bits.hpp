// bits.hpp -- RBSP bit reader for the host-side syntax decoder.
//
// Behavioural contract follows the reference reader (h264bsd_stream.c:73-243,
// h264bsd_vlc.c:105-393): bits past the end of the NAL read as zero, and any read that
// crosses the end is an error (END_OF_STREAM there, `ok == false` here).
#pragma once
#include <cstdint>
#include <cstddef>

namespace b200 {

class BitReader {
public:
    BitReader() = default;
    BitReader(const uint8_t *data, size_t size) : p_(data), size_(size), bitsTotal_(size * 8) {}

    const uint8_t *data() const { return p_; }
    size_t size() const { return size_; }
    uint64_t pos() const { return pos_; }
    void seek(uint64_t bitpos) { pos_ = bitpos; }
    bool byteAligned() const { return (pos_ & 7) == 0; }
    uint64_t bitsLeft() const { return pos_ >= bitsTotal_ ? 0 : bitsTotal_ - pos_; }

    // next 32 bits, MSB first, zero padded past the end (h264bsdShowBits32)
    uint32_t show32() const {
        uint64_t byte = pos_ >> 3;
        unsigned sh = (unsigned)(pos_ & 7);
        uint64_t v = 0;
        if (byte + 8 <= size_) {
            __builtin_memcpy(&v, p_ + byte, 8);     // one unaligned load, bytes to MSB-first order
#if __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
            v = __builtin_bswap64(v);
#endif
        } else {
            for (unsigned i = 0; i < 8; i++) {
                uint64_t b = (byte + i < size_) ? p_[byte + i] : 0;
                v |= b << (56 - 8 * i);
            }
        }
        return (uint32_t)((v << sh) >> 32);
    }
    uint32_t show(unsigned n) const { return n ? show32() >> (32 - n) : 0; }
    // the stream from bit position `bitpos` on, MSB first: at least 57 bits are stream bits (zero past the end), the rest 0
    uint64_t window(uint64_t bitpos) const {
        const uint64_t byte = bitpos >> 3;
        uint64_t v = 0;
        if (byte + 8 <= size_) {
            __builtin_memcpy(&v, p_ + byte, 8);
#if __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
            v = __builtin_bswap64(v);
#endif
        } else {
            for (unsigned i = 0; i < 8; i++) {
                uint64_t b = (byte + i < size_) ? p_[byte + i] : 0;
                v |= b << (56 - 8 * i);
            }
        }
        return v << (bitpos & 7);
    }
    uint64_t bitsTotal() const { return bitsTotal_; }

    // h264bsdFlushBits: returns false when the read crossed the end of the NAL
    bool skip(unsigned n) {
        pos_ += n;
        return pos_ <= bitsTotal_;
    }
    // h264bsdGetBits (n < 32)
    bool get(unsigned n, uint32_t &out) {
        out = show(n);
        return skip(n);
    }
    bool get1(uint32_t &out) { return get(1, out); }

    // ue(v) -- h264bsdDecodeExpGolombUnsigned
    bool ue(uint32_t &val) {
        uint32_t bits = show32();
        if (bits >= 0x80000000u) {
            val = 0;
            return skip(1);
        }
        unsigned zeros = (unsigned)__builtin_clz(bits | 1u);
        if (bits == 0) zeros = 32;
        if (zeros >= 32) {
            // 32 zero bits: only the 2^32-1 / 2^32 corner codes exist (vlc.c:163-185); both are
            // outside every syntax element's legal range -> error.
            skip(32);
            val = 0xFFFFFFFFu;
            return false;
        }
        if (zeros <= 15) {
            uint32_t code = bits >> (31 - 2 * zeros);  // 2*zeros+1 bits
            val = code - 1;
            return skip(2 * zeros + 1);
        }
        if (!skip(zeros + 1)) return false;
        uint32_t suffix = 0;
        if (!get(zeros, suffix)) return false;
        val = ((1u << zeros) - 1) + suffix;
        return true;
    }
    // se(v) -- h264bsdDecodeExpGolombSigned
    bool se(int32_t &val) {
        uint32_t k;
        if (!ue(k)) return false;
        val = (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)((k + 1) >> 1);
        return true;
    }
    // te(v) -- h264bsdDecodeExpGolombTruncated
    bool te(uint32_t &val, bool rangeGreaterThanOne) {
        if (rangeGreaterThanOne) return ue(val);
        uint32_t b;
        if (!get1(b)) return false;
        val = b ^ 1u;
        return true;
    }

    // more_rbsp_data() -- h264bsdMoreRbspData (h264bsd_util.c)
    bool moreRbspData() const {
        uint64_t left = bitsLeft();
        if (left == 0) return false;
        if (left > 8) return true;
        return (show32() >> (32 - left)) != (1u << (left - 1));
    }
    // rbsp_trailing_bits() -- h264bsdRbspTrailingBits
    bool trailingBits() {
        unsigned n = 8 - (unsigned)(pos_ & 7);
        uint32_t v;
        if (!get(n, v)) return false;
        return v == (1u << (n - 1));
    }

private:
    const uint8_t *p_ = nullptr;
    size_t size_ = 0;
    uint64_t bitsTotal_ = 0;
    uint64_t pos_ = 0;
};

}  // namespace b200
