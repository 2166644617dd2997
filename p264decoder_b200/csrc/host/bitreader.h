// Big-endian bit reader over an RBSP payload (emulation prevention already removed).
// Semantics follow the reference reader's observable behaviour (core/bs.h:50-174):
// reads past the end return zero bits and eof() turns true once the byte pointer
// reaches the end.
#pragma once
#include <cstdint>
#include <cstddef>

namespace p264b200 {

class BitReader {
public:
    BitReader(const uint8_t *data, size_t size) : p_(data), size_bits_(size * 8), pos_(0) {}

    bool eof() const { return pos_ >= size_bits_; }
    size_t pos() const { return pos_; }

    // peek up to 25 bits without consuming.  A 64-bit window of the stream is cached (one unaligned big-endian load per
    // ~5 reads instead of one per read: the block reader peeks and skips a few bits at a time)
    uint32_t show(int n) const {
        if (n <= 0) return 0;
        size_t off = pos_ - wpos_;          // wraps to a huge value when the window lies ahead of pos_
        if (off + (size_t)n > 64) {
            refill();
            off = pos_ - wpos_;
        }
        return (uint32_t)((w_ << off) >> (64 - n));
    }
    void skip(int n) { pos_ += n; }
    uint32_t read(int n) {
        uint32_t v = 0;
        while (n > 24) {  // keep show() within its 25-bit window
            v = (v << 16) | show(16);
            skip(16);
            n -= 16;
        }
        if (n > 0) {
            v = (n == 32 ? 0 : (v << n)) | show(n);
            skip(n);
        }
        return v;
    }
    uint32_t read1() { return read(1); }

    // Exp-Golomb ue(v); the leading-zero scan is capped at 32 like core/bs.h:143-152
    int ue() {
        // codes of up to 25 bits (<= 12 leading zeros) straight from one window; longer / truncated ones bit by bit
        const uint32_t w = show(25);
        if (w >> 12) {
            const int zeros = __builtin_clz(w) - 7, len = 2 * zeros + 1;
            skip(len);
            return (int)((w >> (25 - len)) - 1);
        }
        int zeros = 0;
        while (!eof() && read1() == 0 && zeros < 32) zeros++;
        if (zeros == 0) return 0;
        if (zeros >= 32) return -1;
        return (int)((1u << zeros) - 1 + read(zeros));
    }
    int se() {
        int v = ue();
        return (v & 1) ? (v + 1) / 2 : -(v / 2);
    }
    // te(v) with range x (core/bs.h:160-171)
    int te(int x) {
        if (x == 1) return 1 - (int)read1();
        if (x > 1) return ue();
        return 0;
    }

private:
    void refill() const {
        const size_t byte = pos_ >> 3, nbytes = size_bits_ >> 3;
        wpos_ = byte << 3;
        if (byte + 8 <= nbytes) {
            uint64_t w;
            __builtin_memcpy(&w, p_ + byte, 8);
            w_ = __builtin_bswap64(w);
        } else {
            // tail of the payload: bytes past the end read as zero (core/bs.h behaviour)
            uint64_t w = 0;
            for (int i = 0; i < 8; i++) w = (w << 8) | (byte + i < nbytes ? p_[byte + i] : 0);
            w_ = w;
        }
    }
    const uint8_t *p_;
    size_t size_bits_;
    size_t pos_;
    mutable uint64_t w_ = 0;
    mutable size_t wpos_ = ~(size_t)0 - 1024;   // bit position of the cached window's first bit (byte aligned); none yet
};

}  // namespace p264b200
