/* Integer math helpers (power-of-two, log, alignment, block splitting). */
#ifndef UCC_MATH_H_
#define UCC_MATH_H_
#include "ucc_compiler_def.h"

#define ucc_div_round_up(_n, _d) (((_n) + (_d) - 1) / (_d))
#define ucc_align_down(_n, _a)   ((_n) - ((_n) % (_a)))
#define ucc_align_up(_n, _a)     ucc_align_down((_n) + (_a) - 1, (_a))
#define ucc_align_down_pow2(_n, _a) ((_n) & ~((__typeof__(_n))(_a) - 1))
#define ucc_align_up_pow2(_n, _a)   ucc_align_down_pow2((_n) + (_a) - 1, (_a))
#define ucc_is_pow2_or_zero(_n) (!((_n) & ((_n) - 1)))
#define ucc_is_pow2(_n) (((_n) > 0) && ucc_is_pow2_or_zero(_n))
#define ucc_padding(_n, _a) ((_a - ((_n) % (_a))) % (_a))

static inline unsigned ucc_ilog2(uint64_t n) { return n ? 63 - __builtin_clzll(n) : 0; }
/* smallest p with radix^p >= size; *full = radix^p */
static inline unsigned ucc_ilog_radix(uint64_t size, unsigned radix, uint64_t *full)
{
    unsigned p = 0; uint64_t f = 1;
    while (f < size) { f *= radix; p++; }
    if (full) *full = f;
    return p;
}
static inline uint64_t ucc_pow_u64(uint64_t b, unsigned e)
{ uint64_t r = 1; while (e--) r *= b; return r; }
static inline unsigned ucc_popcount64(uint64_t v) { return (unsigned)__builtin_popcountll(v); }
static inline unsigned ucc_ffs64(uint64_t v) { return (unsigned)__builtin_ctzll(v); }

/* Split `total` items in `n` nearly-equal blocks: block i gets
 * total/n (+1 for the first total%n blocks). */
static inline size_t ucc_buffer_block_count(size_t total, unsigned n, unsigned i)
{ size_t b = total / n, r = total % n; return i < r ? b + 1 : b; }
static inline size_t ucc_buffer_block_offset(size_t total, unsigned n, unsigned i)
{ size_t b = total / n, r = total % n; return i * b + (i < r ? i : r); }
/* same but every block boundary aligned to `align` items */
static inline size_t ucc_buffer_block_count_aligned(size_t total, unsigned n, unsigned i, size_t align)
{
    size_t per = ucc_align_up(ucc_div_round_up(total, n), align);
    size_t off = (size_t)i * per;
    if (off >= total) return 0;
    return ucc_min(per, total - off);
}
static inline size_t ucc_buffer_block_offset_aligned(size_t total, unsigned n, unsigned i, size_t align)
{
    size_t per = ucc_align_up(ucc_div_round_up(total, n), align);
    size_t off = (size_t)i * per;
    return ucc_min(off, total);
}
#define PTR_OFFSET(_p, _off) ((void *)((char *)(_p) + (ptrdiff_t)(_off)))
#endif
