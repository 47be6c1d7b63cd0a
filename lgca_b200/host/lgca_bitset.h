// Bit-field with the interface of the reference's lgca::Bitset (src/lgca_bitset.h:31-245): uint8 blocks,
// LSB-first bit addressing (bit i lives in block i/8 at position i%8), a proxy for b[i] = x, whole-block
// access with b(i), ptr(), copy(), resize(), fill_random().  Own implementation; the storage can come from a
// custom allocator so that the B200 backend keeps its host mirrors in pinned memory.
#ifndef LGCA_B200_HOST_BITSET_H_
#define LGCA_B200_HOST_BITSET_H_

#include <cstring>
#include <iostream>
#include <limits>

#include "lgca_common.h"

namespace lgca {

class Bitset {
public:
    using Block = uint8_t;
    static constexpr Block BITS_PER_BLOCK = std::numeric_limits<Block>::digits;

    typedef void* (*AllocFn)(size_t bytes);
    typedef void (*FreeFn)(void* p);

    // proxy so that `b[i] = x`, `b[i] ^= x`, `bool(b[i])` work on single bits
    class reference {
    public:
        reference(Block* blk, Block mask) : blk_(blk), mask_(mask) {}
        operator bool() const { return (*blk_ & mask_) != 0; }
        bool operator~() const { return (*blk_ & mask_) == 0; }
        reference& operator=(bool x) { if (x) *blk_ |= mask_; else *blk_ &= (Block)~mask_; return *this; }
        reference& operator=(const reference& o) { return *this = bool(o); }
        reference& operator|=(bool x) { if (x) *blk_ |= mask_; return *this; }
        reference& operator&=(bool x) { if (!x) *blk_ &= (Block)~mask_; return *this; }
        reference& operator^=(bool x) { if (x) *blk_ ^= mask_; return *this; }
        reference& flip() { *blk_ ^= mask_; return *this; }
    private:
        Block* blk_;
        Block  mask_;
    };

    Bitset() {}
    explicit Bitset(size_t nbits) { resize(nbits); }
    Bitset(const Bitset&) = delete;
    Bitset& operator=(const Bitset&) = delete;
    ~Bitset() { release(); }

    // storage hooks (must be set before resize); default calloc/free
    void set_allocator(AllocFn a, FreeFn f) { alloc_ = a; free_ = f; }

    void resize(size_t nbits)
    {
        release();
        nbits_   = nbits;
        nblocks_ = (nbits + BITS_PER_BLOCK - 1) / BITS_PER_BLOCK;
        const size_t bytes = nblocks_ ? nblocks_ : 1;
        bits_ = static_cast<Block*>(alloc_ ? alloc_(bytes) : std::malloc(bytes));
        if (!bits_) { std::fprintf(stderr, "ERROR in Bitset::resize(): out of memory.\n"); std::abort(); }
        std::memset(bits_, 0, bytes);
    }

    bool operator[](size_t pos) const { assert(pos < nbits_); return (bits_[pos >> 3] >> (pos & 7)) & 1u; }
    reference operator[](size_t pos) { assert(pos < nbits_); return reference(bits_ + (pos >> 3), Block(1u << (pos & 7))); }
    Block operator()(size_t blk) const { assert(blk < nblocks_); return bits_[blk]; }
    Block& operator()(size_t blk) { assert(blk < nblocks_); return bits_[blk]; }

    void set(size_t pos, bool v = true) { (*this)[pos] = v; }
    void reset(size_t pos) { (*this)[pos] = false; }
    void reset() { std::memset(bits_, 0, nblocks_); }
    void flip(size_t pos) { (*this)[pos].flip(); }

    size_t size() const { return nbits_; }
    size_t num_blocks() const { return nblocks_; }
    size_t count() const
    {
        size_t c = 0;
        for (size_t i = 0; i < nblocks_; ++i) c += (size_t)__builtin_popcount(bits_[i]);
        return c;
    }
    void print() const
    {
        for (size_t i = 0; i < nbits_; ++i) std::cout << (*this)[i] << " ";
        std::cout << std::endl;
    }
    void copy(const Bitset& other)
    {
        assert(other.nbits_ == nbits_);
        std::memcpy(bits_, other.bits_, nblocks_);
    }
    Block* ptr() { return bits_; }
    const Block* ptr() const { return bits_; }

    // bit i = rand() % 2, ascending i -- the chirality field's place in the process-wide rand() stream
    // (reference: src/lgca_bitset.h:220-224)
    void fill_random()
    {
        for (size_t i = 0; i < nbits_; ++i) (*this)[i] = (std::rand() % 2) != 0;
    }

private:
    void release()
    {
        if (bits_) { if (free_) free_(bits_); else std::free(bits_); }
        bits_ = nullptr;
        nbits_ = nblocks_ = 0;
    }
    Block*  bits_ = nullptr;
    size_t  nbits_ = 0, nblocks_ = 0;
    AllocFn alloc_ = nullptr;
    FreeFn  free_ = nullptr;
};

} // namespace lgca

#endif
