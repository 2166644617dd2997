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

    // peek up to 25 bits without consuming
    uint32_t show(int n) const {
        if (n <= 0) return 0;
        size_t byte = pos_ >> 3;
        int sh = pos_ & 7;
        size_t nbytes = size_bits_ >> 3;
        if (byte + 8 <= nbytes) {
            // fast path: one unaligned big-endian 64-bit load (the host parser is the bottleneck of real-bitstream decode)
            uint64_t w;
            __builtin_memcpy(&w, p_ + byte, 8);
            w = __builtin_bswap64(w);
            return (uint32_t)((w << sh) >> (64 - n));
        }
        uint64_t w = 0;
        for (int i = 0; i < 5; i++) w = (w << 8) | (byte + i < nbytes ? p_[byte + i] : 0);
        return (uint32_t)((w >> (40 - sh - n)) & ((1ull << n) - 1));
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
    const uint8_t *p_;
    size_t size_bits_;
    size_t pos_;
};

}  // namespace p264b200
