/* simmr_types.h — plain-data structures shared by the host model and the CUDA
 * kernels of libsimmr_b200.  Nothing here depends on CUDA headers. */
#pragma once
#include <stdint.h>

namespace simmr {

/* header token kinds: the placeholders of fastq.rs:34-57 */
enum : uint8_t { TOK_LIT = 0, TOK_READ_ID = 1, TOK_START = 2, TOK_END = 3, TOK_RC = 4, TOK_PAIR = 5 };

struct HdrTok {        /* one piece of a compiled per-contig header template */
    uint32_t off;      /* TOK_LIT: offset into the literal pool              */
    uint16_t len;      /* TOK_LIT: byte length                               */
    uint8_t  type;
    uint8_t  pad_;
};
struct ContigHdr {     /* per (contig, mate): token span + what is needed for H; index 2*contig + mate */
    uint32_t tok_off;
    uint16_t ntok;
    uint16_t lit_bytes;
    uint8_t  n_id, n_sp, n_ep, n_rc, n_pair, pad_[3];
};

/* per-position alias tables for qualities */
enum : uint32_t { PT_CONST = 0,   /* every base has quality const_q, nothing to sample */
                  PT_Q32 = 3 };   /* thr16<<16 | alias<<8 | self per column, characters as printed; the low cb bits of a draw's word
                                     pick the column, self iff the draw's level < entry; substitutions are a separate sparse
                                     process (gap / qsub tables).  Kinds 1 and 2 (joint and top-bit tables) are retired. */
struct PhredTab {
    uint32_t kind, const_q;
    uint32_t cb;           /* log2(columns)                                   */
    uint32_t npos;         /* positions; base i uses min(i, npos-1)           */
    const uint32_t *entries;   /* device or host pointer                      */
    /* mutating profile: gap[256] = bases until the next substitution (255 = none within 255),
     * qsub[1 << cb] = quality given a substitution; has_sub == 0 when the profile never substitutes */
    const uint32_t *gap, *qsub;
    uint32_t has_sub, pad_;
};
constexpr uint32_t GAP_COLS = 256, GAP_NONE = 255;
/* length sampler: constant, alias over bins (entry = thr20<<12 | alias12) or
 * Marsaglia-Tsang gamma */
enum : uint32_t { LEN_CONST = 0, LEN_ALIAS = 1, LEN_GAMMA = 2 };
struct LenTab {
    uint32_t kind, value, cb, pad_;
    const uint32_t *entries, *lo, *hi;
};

struct GammaConst { double d, c, scale; };

/* k-mer error injection tables of a custom-long profile (ErrModel in host_model.hpp), device pointers */
struct ErrTab {
    uint32_t on, k, n_ref, pad_;
    const uint32_t *direct;        /* 2 words per A,C,G,T k-mer (2-bit code): keep threshold, row + 1 (0: not observed) */
    const uint32_t *ref, *keep_thr, *row, *cb, *entries, *alts;
    /* the same tables as the 2-bit walk (reads of A,C,G,T only) wants them — two dependent loads per error instead of five:
     *   direct2[4^k]  {keep_thr, (first entry of the k-mer's alias row + 1) | cb << 28}, 0 in the second word: k-mer not observed
     *   ent4[entry]   {thr20, the column's alternate, its alias's alternate, 0}; an alternate = its bases (A,C,G,T = 0..3, N
     *                 digits dropped) in the low bits, their count << 28 */
    const uint32_t *direct2, *ent4;
};   /* perfect_long.rs:48-54 via rand_distr gamma.rs */

/* which reference arithmetic the replay path applies per base */
enum : uint32_t { RP_CONST = 0, RP_NORMAL_FLOOR = 1, RP_ACCURACY = 2, RP_ALIAS = 3 };

}  // namespace simmr
