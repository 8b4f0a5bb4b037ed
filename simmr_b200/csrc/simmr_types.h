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
struct ContigHdr {     /* per contig: token span + what is needed for H      */
    uint32_t tok_off;
    uint16_t ntok;
    uint16_t lit_bytes;
    uint8_t  n_id, n_sp, n_ep, n_rc, n_pair, pad_[3];
};

/* per-position quantised alias tables: entry = thr22<<10 | alias10 */
struct PhredTab {
    uint32_t constant;     /* 1: every base has quality const_q, no table     */
    uint32_t const_q;
    uint32_t cb;           /* log2(columns)                                   */
    uint32_t npos;         /* positions; base i uses min(i, npos-1)           */
    uint32_t qlo;          /* quality of outcome 0                            */
    uint32_t mshift;       /* 2: outcome = q<<2 | mutation, 0: outcome = q    */
    const uint32_t *entries;   /* device or host pointer, npos<<cb words      */
};
/* length sampler: constant, alias over bins (entry = thr20<<12 | alias12) or
 * Marsaglia-Tsang gamma */
enum : uint32_t { LEN_CONST = 0, LEN_ALIAS = 1, LEN_GAMMA = 2 };
struct LenTab {
    uint32_t kind, value, cb, pad_;
    const uint32_t *entries, *lo, *hi;
};

struct GammaConst { double d, c, scale; };   /* perfect_long.rs:48-54 via rand_distr gamma.rs */

/* which reference arithmetic the replay path applies per base */
enum : uint32_t { RP_CONST = 0, RP_NORMAL_FLOOR = 1, RP_ACCURACY = 2, RP_ALIAS = 3 };

}  // namespace simmr
