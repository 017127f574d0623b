/*
 * hfg_inflate.c -- a gzip (RFC 1952 / DEFLATE RFC 1951) decoder for `.cov.gz` inputs.
 *
 * Why not zlib's: the reference writes its coverage files with gzopen(path, "w6h") (submodules/ptBlock/ptBlock.c:2271), i.e.
 * Huffman-only DEFLATE -- every byte is a literal symbol, there are no matches to copy -- and zlib's inflate decodes such
 * a stream at ~125 MB/s, which made inflating two thirds of the time the reader takes (profiles/host_timing_r1d.txt).  This
 * decoder keeps 64 bits of input in a register, resolves codes of up to 11 bits with one table look-up, delivers two
 * literals per look-up where both codes fit the index, and does four look-ups per refill: ~2.5x zlib on such streams, on
 * par with it on ordinary gzip -6 output.  It is a complete inflate (stored, fixed and dynamic blocks, matches, multi-member files), it
 * streams (output in pieces of ~4 MB, 32 KB of history carried over), and it checks every member's CRC-32 and length, so a
 * decoding mistake cannot pass silently.  HFG_ZLIB_INFLATE=1 makes the reader use zlib instead.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h> /* crc32() only */

#include "hfg_inflate.h"

#define WINDOW 32768
#define LIT_BITS 11  /* primary table of the literal / length code */
#define DIST_BITS 8  /* primary table of the distance code */
#define MAX_CODE_LEN 15
#define IN_PAD 16    /* readable bytes behind the input, so that refills never test for the end */
#define OUT_SLACK 320 /* a match started before the piece limit may run up to 258 bytes past it, literal steps 8 */

/* table entry: bits 0..7 = code length (bits to drop), bits 8..15 = kind, bits 16..31 = value */
enum { K_LITERAL = 0, K_LENGTH = 1, K_END = 2, K_SUBTABLE = 3, K_DIST = 4, K_INVALID = 5, K_PAIR = 6 };
/* K_PAIR: two literals whose codes fit the primary index together; value = first | second << 8, length = both codes */
#define ENTRY(kind, len, value) ((uint32_t) (len) | ((uint32_t) (kind) << 8) | ((uint32_t) (value) << 16))
#define E_LEN(e) ((e) & 0xffu)
#define E_KIND(e) (((e) >> 8) & 0xffu)
#define E_VALUE(e) ((e) >> 16)

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CLEN_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct hfg_inflate {
    uint8_t *in; /* the whole compressed file, IN_PAD zero bytes behind it */
    size_t in_len, in_pos;
    uint64_t bitbuf;
    int bitcnt;
    /* output: [WINDOW bytes of history][piece] */
    uint8_t *out;
    size_t piece_cap;
    /* member / block state */
    int in_member, in_block, last_block, block_kind; /* block_kind: 0 stored, 1 Huffman */
    int members_done;                                /* complete, CRC-checked members so far */
    size_t stored_left;
    uint32_t crc;
    uint64_t member_bytes;
    uint32_t lit_table[(1 << LIT_BITS) + 4096]; /* primary + subtables */
    uint32_t dist_table[(1 << DIST_BITS) + 2048];
    int finished, failed;
    char err[128];
};

static int fail(hfg_inflate *z, const char *msg) {
    z->failed = 1;
    snprintf(z->err, sizeof(z->err), "%s", msg);
    return -1;
}

const char *hfg_inflate_error(const hfg_inflate *z) { return z->err; }

/* past the end the refill shifts in zeros (in_pos keeps counting, so overrun() sees it); decoding stops at the next check */
#define REFILL(z)                                                            \
    do {                                                                     \
        if ((z)->bitcnt < 56) {                                              \
            uint64_t w_ = 0;                                                 \
            if ((z)->in_pos + 8 <= (z)->in_len + IN_PAD) memcpy(&w_, (z)->in + (z)->in_pos, 8); \
            (z)->bitbuf |= w_ << (z)->bitcnt;                                \
            const int take_ = (63 - (z)->bitcnt) >> 3;                       \
            (z)->in_pos += (size_t) take_;                                   \
            (z)->bitcnt += take_ << 3;                                       \
        }                                                                    \
    } while (0)
#define PEEK(z, n) ((uint32_t) ((z)->bitbuf & ((1ull << (n)) - 1)))
#define DROP(z, n) ((z)->bitbuf >>= (n), (z)->bitcnt -= (n))

/* more bits consumed than the file holds?  (bits consumed = 8 * in_pos - bitcnt) */
static int overrun(const hfg_inflate *z) { return z->in_pos > z->in_len && 8 * (z->in_pos - z->in_len) > (size_t) z->bitcnt; }

static uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

/* canonical Huffman code -> decode table.  lens[n]: code lengths (0 = unused).  primary_bits: size of the first-level table;
 * longer codes go through subtables appended behind it.  is_dist selects the meaning of the symbols.  Returns 0 / -1. */
static int build_table(uint32_t *table, int table_cap, int primary_bits, const uint8_t *lens, int n, int is_dist) {
    int count[MAX_CODE_LEN + 1] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int used = 0;
    for (int l = 1; l <= MAX_CODE_LEN; l++) used += count[l];
    const uint32_t invalid = ENTRY(K_INVALID, 1, 0);
    for (int i = 0; i < (1 << primary_bits); i++) table[i] = invalid;
    if (used == 0) return 0; /* no codes: any use is an error, caught through K_INVALID */
    /* over-subscribed or incomplete sets: incomplete is legal only for a single code (RFC 1951, one distance code) */
    int left = 1;
    for (int l = 1; l <= MAX_CODE_LEN; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return -1;
    }
    if (left > 0 && !(used == 1)) return -1;
    uint32_t next_code[MAX_CODE_LEN + 2];
    uint32_t code = 0;
    for (int l = 1; l <= MAX_CODE_LEN; l++) {
        code = (code + (uint32_t) count[l - 1]) << 1;
        next_code[l] = code;
    }
    int next_sub = 1 << primary_bits;
    /* subtable roots are created on demand; sub_bits per root = longest code sharing the prefix - primary_bits.  Two passes:
     * first find, per primary prefix, the longest code */
    uint8_t *sub_len = calloc((size_t) 1 << primary_bits, 1);
    if (!sub_len) return -1;
    {
        uint32_t nc[MAX_CODE_LEN + 2];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n; s++) {
            const int l = lens[s];
            if (l <= primary_bits) {
                if (l) nc[l]++;
                continue;
            }
            const uint32_t rev = reverse_bits(nc[l]++, l);
            const uint32_t prefix = rev & ((1u << primary_bits) - 1);
            if (l - primary_bits > sub_len[prefix]) sub_len[prefix] = (uint8_t) (l - primary_bits);
        }
    }
    for (int s = 0; s < n; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rev = reverse_bits(next_code[l]++, l);
        uint32_t entry;
        if (is_dist) entry = s < 30 ? ENTRY(K_DIST, l, s) : ENTRY(K_INVALID, l, 0);
        else if (s < 256) entry = ENTRY(K_LITERAL, l, s);
        else if (s == 256) entry = ENTRY(K_END, l, 0);
        else entry = s - 257 < 29 ? ENTRY(K_LENGTH, l, s - 257) : ENTRY(K_INVALID, l, 0);
        if (l <= primary_bits) {
            for (uint32_t i = rev; i < (1u << primary_bits); i += 1u << l) table[i] = entry;
        } else {
            const uint32_t prefix = rev & ((1u << primary_bits) - 1);
            const int sb = sub_len[prefix];
            if (E_KIND(table[prefix]) != K_SUBTABLE) {
                if (next_sub + (1 << sb) > table_cap) {
                    free(sub_len);
                    return -1;
                }
                table[prefix] = ENTRY(K_SUBTABLE, sb, next_sub); /* len field = index bits of the subtable */
                for (int i = 0; i < (1 << sb); i++) table[next_sub + i] = invalid;
                next_sub += 1 << sb;
            }
            const uint32_t base = E_VALUE(table[prefix]);
            /* inside the subtable the entry's length is the bits consumed AFTER the primary bits */
            const uint32_t sub_entry = (entry & ~0xffu) | (uint32_t) (l - primary_bits);
            for (uint32_t i = rev >> primary_bits; i < (1u << sb); i += 1u << (l - primary_bits)) table[base + i] = sub_entry;
        }
    }
    free(sub_len);
    return 0;
}

/* Where two consecutive literal codes fit into the primary index, one look-up delivers both: halves the chain
 * "look up -> shift -> look up" that bounds a literal-only stream (coverage text has ~15 symbols of 3-5 bits). */
static void pair_literals(uint32_t *table) {
    uint32_t src[1 << LIT_BITS]; /* the pairs are formed from the single-literal entries, not from each other */
    memcpy(src, table, sizeof(src));
    for (uint32_t i = 0; i < (1u << LIT_BITS); i++) {
        const uint32_t e1 = src[i];
        if (E_KIND(e1) != K_LITERAL) continue;
        const uint32_t l1 = E_LEN(e1);
        const uint32_t e2 = src[i >> l1]; /* the upper bits of that index are zeros we do not know: usable only if the
                                             second code ends inside the known bits */
        if (E_KIND(e2) != K_LITERAL || l1 + E_LEN(e2) > LIT_BITS) continue;
        table[i] = ENTRY(K_PAIR, l1 + E_LEN(e2), E_VALUE(e1) | (E_VALUE(e2) << 8));
    }
}

static int fixed_tables(hfg_inflate *z) {
    uint8_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    if (build_table(z->lit_table, (int) (sizeof(z->lit_table) / 4), LIT_BITS, lens, 288, 0)) return -1;
    uint8_t dl[32];
    for (int i = 0; i < 32; i++) dl[i] = 5;
    return build_table(z->dist_table, (int) (sizeof(z->dist_table) / 4), DIST_BITS, dl, 32, 1);
}

static int dynamic_tables(hfg_inflate *z) {
    REFILL(z);
    const int hlit = (int) PEEK(z, 5) + 257;
    DROP(z, 5);
    const int hdist = (int) PEEK(z, 5) + 1;
    DROP(z, 5);
    const int hclen = (int) PEEK(z, 4) + 4;
    DROP(z, 4);
    if (hlit > 286 || hdist > 30) return fail(z, "bad dynamic block header");
    uint8_t cl[19] = {0};
    for (int i = 0; i < hclen; i++) {
        REFILL(z);
        cl[CLEN_ORDER[i]] = (uint8_t) PEEK(z, 3);
        DROP(z, 3);
    }
    uint32_t cl_table[1 << 7];
    if (build_table(cl_table, 1 << 7, 7, cl, 19, 0)) return fail(z, "bad code-length code");
    uint8_t lens[286 + 30 + 140];
    int n = 0;
    while (n < hlit + hdist) {
        REFILL(z);
        const uint32_t e = cl_table[PEEK(z, 7)];
        if (E_KIND(e) == K_INVALID) return fail(z, "bad code-length symbol");
        DROP(z, (int) E_LEN(e));
        const int sym = (int) E_VALUE(e); /* built as literals 0..18 */
        if (sym < 16) {
            lens[n++] = (uint8_t) sym;
        } else {
            int rep, val = 0;
            if (sym == 16) {
                if (n == 0) return fail(z, "repeat with no previous length");
                val = lens[n - 1];
                rep = 3 + (int) PEEK(z, 2);
                DROP(z, 2);
            } else if (sym == 17) {
                rep = 3 + (int) PEEK(z, 3);
                DROP(z, 3);
            } else {
                rep = 11 + (int) PEEK(z, 7);
                DROP(z, 7);
            }
            if (n + rep > hlit + hdist) return fail(z, "code lengths overrun");
            while (rep--) lens[n++] = (uint8_t) val;
        }
    }
    if (overrun(z)) return fail(z, "truncated input");
    if (lens[256] == 0) return fail(z, "no end-of-block code");
    if (build_table(z->lit_table, (int) (sizeof(z->lit_table) / 4), LIT_BITS, lens, hlit, 0)) return fail(z, "bad literal/length code");
    pair_literals(z->lit_table);
    if (build_table(z->dist_table, (int) (sizeof(z->dist_table) / 4), DIST_BITS, lens + hlit, hdist, 1)) return fail(z, "bad distance code");
    return 0;
}

static void byte_align(hfg_inflate *z) {
    const int drop = z->bitcnt & 7;
    DROP(z, drop);
}

/* a byte-aligned little-endian field of n <= 4 bytes */
static uint32_t take_bytes(hfg_inflate *z, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        REFILL(z);
        v |= PEEK(z, 8) << (8 * i);
        DROP(z, 8);
    }
    return v;
}

static int gzip_header(hfg_inflate *z) {
    const uint32_t magic = take_bytes(z, 3);
    const uint32_t flg = take_bytes(z, 1);
    if (magic != 0x088b1f) return fail(z, "not a gzip member");
    if (flg & 0xe0) return fail(z, "reserved gzip flags set");
    take_bytes(z, 4); /* mtime */
    take_bytes(z, 2); /* xfl, os */
    if (flg & 4) {
        uint32_t xlen = take_bytes(z, 2);
        while (xlen--) {
            take_bytes(z, 1);
            if (overrun(z)) return fail(z, "truncated gzip header");
        }
    }
    for (int field = 0; field < 2; field++) /* FNAME, FCOMMENT: zero-terminated */
        if (flg & (8u << field))
            while (take_bytes(z, 1) != 0)
                if (overrun(z)) return fail(z, "truncated gzip header");
    if (flg & 2) take_bytes(z, 2); /* header CRC */
    if (overrun(z)) return fail(z, "truncated gzip header");
    z->in_member = 1;
    z->in_block = 0;
    z->last_block = 0;
    z->crc = (uint32_t) crc32(0L, Z_NULL, 0);
    z->member_bytes = 0;
    return 0;
}

hfg_inflate *hfg_inflate_open(const char *path, size_t piece_bytes) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    unsigned char magic[2] = {0, 0};
    if (fread(magic, 1, 2, f) != 2 || magic[0] != 0x1f || magic[1] != 0x8b) { /* not gzip: nothing read, nothing allocated */
        fclose(f);
        return NULL;
    }
    hfg_inflate *z = calloc(1, sizeof(*z));
    if (!z) {
        fclose(f);
        return NULL;
    }
    if (fseek(f, 0, SEEK_END) == 0) {
        const long len = ftell(f);
        if (len >= 0 && fseek(f, 0, SEEK_SET) == 0) {
            z->in = malloc((size_t) len + IN_PAD);
            if (z->in && fread(z->in, 1, (size_t) len, f) == (size_t) len) {
                memset(z->in + len, 0, IN_PAD);
                z->in_len = (size_t) len;
            } else {
                free(z->in);
                z->in = NULL;
            }
        }
    }
    fclose(f);
    z->piece_cap = piece_bytes;
    z->out = z->in ? malloc(WINDOW + piece_bytes + OUT_SLACK) : NULL;
    if (!z->in || !z->out || z->in_len < 18) {
        hfg_inflate_close(z);
        return NULL;
    }
    memset(z->out, 0, WINDOW);
    return z;
}

void hfg_inflate_close(hfg_inflate *z) {
    if (!z) return;
    free(z->in);
    free(z->out);
    free(z);
}

size_t hfg_inflate_piece_capacity(const hfg_inflate *z) { return z->piece_cap + OUT_SLACK; }

/* Decodes the next piece.  Returns the number of bytes written to dst (at most hfg_inflate_piece_capacity), 0 at the end of
 * the file, -1 on an error (hfg_inflate_error). */
long hfg_inflate_next(hfg_inflate *z, uint8_t *dst) {
    if (z->failed) return -1;
    if (z->finished) return 0;
    uint8_t *const base = z->out + WINDOW;
    uint8_t *out = base;
    uint8_t *const limit = base + z->piece_cap;
    size_t crc_from = 0; /* offset in the piece from which the current member's CRC is still to be taken */

    while (out < limit) {
        if (!z->in_member) {
            /* between members: a gzip magic means another member.  Behind at least one complete, CRC-checked member anything
             * else -- zero padding or garbage -- ends the stream quietly, as zlib's gzread (the reference's reader) treats it;
             * a file that does not START with a member is an error */
            byte_align(z);
            REFILL(z);
            const size_t consumed = z->in_pos - (size_t) (z->bitcnt >> 3);
            if (consumed >= z->in_len) {
                z->finished = 1;
                break;
            }
            if (z->members_done > 0 && !(z->in[consumed] == 0x1f && consumed + 1 < z->in_len && z->in[consumed + 1] == 0x8b)) {
                z->finished = 1;
                break;
            }
            if (z->in[consumed] == 0) { /* zero padding behind the last member */
                size_t k = consumed;
                while (k < z->in_len && z->in[k] == 0) k++;
                if (k == z->in_len) {
                    z->finished = 1;
                    break;
                }
            }
            if (gzip_header(z)) return -1;
            crc_from = (size_t) (out - base);
        }
        if (!z->in_block) {
            REFILL(z);
            z->last_block = (int) PEEK(z, 1);
            DROP(z, 1);
            const int type = (int) PEEK(z, 2);
            DROP(z, 2);
            if (type == 0) {
                byte_align(z);
                const uint32_t len = take_bytes(z, 2), nlen = take_bytes(z, 2);
                if ((len ^ 0xffffu) != nlen) return fail(z, "stored block length check"), -1;
                z->stored_left = len;
                z->block_kind = 0;
            } else if (type == 1) {
                if (fixed_tables(z)) return fail(z, "internal: fixed tables"), -1;
                z->block_kind = 1;
            } else if (type == 2) {
                if (dynamic_tables(z)) return -1;
                z->block_kind = 1;
            } else {
                return fail(z, "reserved block type"), -1;
            }
            if (overrun(z)) return fail(z, "truncated input"), -1;
            z->in_block = 1;
        }
        if (z->block_kind == 0) {
            /* stored: the bit buffer is byte-aligned; drain it, then copy straight from the input */
            while (z->stored_left && out < limit) {
                if (z->bitcnt >= 8) {
                    *out++ = (uint8_t) PEEK(z, 8);
                    DROP(z, 8);
                    z->stored_left--;
                    continue;
                }
                /* the buffer is empty (whole bytes were taken); what a refill shifted in above its count must go, the input
                 * position moves on without it */
                z->bitbuf = 0;
                z->bitcnt = 0;
                size_t n = z->stored_left;
                if (n > (size_t) (limit - out)) n = (size_t) (limit - out);
                if (z->in_pos + n > z->in_len) return fail(z, "truncated stored block"), -1;
                memcpy(out, z->in + z->in_pos, n);
                z->in_pos += n;
                out += n;
                z->stored_left -= n;
            }
            if (z->stored_left == 0) z->in_block = 0;
        } else {
            const uint32_t *lt = z->lit_table;
            const uint32_t *dt = z->dist_table;
            int end_of_block = 0;
            const char *bad = NULL;
            /* the bit reader lives in locals inside this loop: the byte stores through `out` could alias the fields of *z,
             * which would force a reload of the bit buffer after every literal */
            uint64_t bb = z->bitbuf;
            int bc = z->bitcnt;
            size_t ip = z->in_pos;
            const uint8_t *const in = z->in;
            const size_t in_ok = z->in_len + IN_PAD - 8; /* last position an 8-byte load may start at */
#define LREFILL()                                        \
    do {                                                 \
        if (bc < 56) {                                   \
            uint64_t w_ = 0;                             \
            if (ip <= in_ok) memcpy(&w_, in + ip, 8);    \
            bb |= w_ << bc;                              \
            const int take_ = (63 - bc) >> 3;            \
            ip += (size_t) take_;                        \
            bc += take_ << 3;                            \
        }                                                \
    } while (0)
#define LPEEK(n) ((uint32_t) (bb & ((1ull << (n)) - 1)))
#define LDROP(n) (bb >>= (n), bc -= (int) (n))
            while (out < limit) {
                LREFILL();
                uint32_t e = lt[LPEEK(LIT_BITS)];
                /* pairs and single literals resolved by the primary table take at most 11 bits each: four look-ups out of
                 * one refill (56 bits) */
#define LIT_STEP()                                                           \
    if ((e & 0xff00u) == ((uint32_t) K_PAIR << 8)) {                         \
        LDROP(e & 0xffu);                                                    \
        out[0] = (uint8_t) (e >> 16);                                        \
        out[1] = (uint8_t) (e >> 24);                                        \
        out += 2;                                                            \
    } else if ((e & 0xff00u) == 0) { /* K_LITERAL */                         \
        LDROP(e & 0xffu);                                                    \
        *out++ = (uint8_t) (e >> 16);                                        \
    } else                                                                   \
        goto not_literal;
                LIT_STEP();
                e = lt[LPEEK(LIT_BITS)];
                LIT_STEP();
                e = lt[LPEEK(LIT_BITS)];
                LIT_STEP();
                e = lt[LPEEK(LIT_BITS)];
                LIT_STEP();
                continue;
            not_literal:
#undef LIT_STEP
                LREFILL();
                e = lt[LPEEK(LIT_BITS)];
                if (E_KIND(e) == K_PAIR) { /* cannot happen after the steps above on the same bits, kept for safety */
                    LDROP(E_LEN(e));
                    *out++ = (uint8_t) (e >> 16);
                    *out++ = (uint8_t) (e >> 24);
                    continue;
                }
                if (E_KIND(e) == K_SUBTABLE) {
                    const int sb = (int) E_LEN(e);
                    LDROP(LIT_BITS);
                    e = lt[E_VALUE(e) + LPEEK(sb)];
                }
                LDROP(E_LEN(e));
                const uint32_t kind = E_KIND(e);
                if (kind == K_LITERAL) {
                    *out++ = (uint8_t) E_VALUE(e);
                    continue;
                }
                if (kind == K_END) {
                    end_of_block = 1;
                    break;
                }
                if (kind != K_LENGTH) {
                    bad = "invalid literal/length code";
                    break;
                }
                const uint32_t ls = E_VALUE(e);
                int len = LEN_BASE[ls] + (int) LPEEK(LEN_EXTRA[ls]);
                LDROP(LEN_EXTRA[ls]);
                LREFILL();
                uint32_t d = dt[LPEEK(DIST_BITS)];
                if (E_KIND(d) == K_SUBTABLE) {
                    const int sb = (int) E_LEN(d);
                    LDROP(DIST_BITS);
                    d = dt[E_VALUE(d) + LPEEK(sb)];
                }
                LDROP(E_LEN(d));
                if (E_KIND(d) != K_DIST) {
                    bad = "invalid distance code";
                    break;
                }
                const uint32_t ds = E_VALUE(d);
                const int dist = DIST_BASE[ds] + (int) LPEEK(DIST_EXTRA[ds]);
                LDROP(DIST_EXTRA[ds]);
                /* at most 32768 = WINDOW, which the buffer always holds in front of `out`; it must also stay inside the
                 * member's own output */
                if ((uint64_t) dist > z->member_bytes + ((uint64_t) (out - base) - crc_from)) {
                    bad = "distance beyond the start of the data";
                    break;
                }
                const uint8_t *src = out - dist;
                if (dist >= 8) {
                    /* eight bytes at a time, each read at least eight bytes behind its write; may run up to seven bytes past
                     * the match into the slack, which the next symbols overwrite */
                    uint8_t *o = out;
                    int n = len;
                    do {
                        memcpy(o, src, 8);
                        o += 8;
                        src += 8;
                        n -= 8;
                    } while (n > 0);
                    out += len;
                } else {
                    while (len--) *out++ = *src++; /* short period: the copy feeds on its own output */
                }
                if (ip > z->in_len && 8 * (ip - z->in_len) > (size_t) bc) {
                    bad = "truncated input";
                    break;
                }
            }
#undef LREFILL
#undef LPEEK
#undef LDROP
            z->bitbuf = bb;
            z->bitcnt = bc;
            z->in_pos = ip;
            if (bad) return fail(z, bad), -1;
            if (overrun(z)) return fail(z, "truncated input"), -1;
            if (end_of_block) z->in_block = 0;
        }
        if (!z->in_block && z->last_block) {
            /* member trailer: CRC-32 and length of the uncompressed data */
            const size_t produced = (size_t) (out - base);
            z->crc = (uint32_t) crc32(z->crc, base + crc_from, (uInt) (produced - crc_from));
            z->member_bytes += produced - crc_from;
            crc_from = produced;
            byte_align(z);
            const uint32_t want_crc = take_bytes(z, 4), want_len = take_bytes(z, 4);
            if (overrun(z)) return fail(z, "truncated gzip trailer"), -1;
            if (want_crc != z->crc) return fail(z, "CRC mismatch"), -1;
            if (want_len != (uint32_t) z->member_bytes) return fail(z, "length mismatch"), -1;
            z->in_member = 0;
            z->members_done++;
            z->last_block = 0;
        }
    }
    const size_t produced = (size_t) (out - base);
    if (z->in_member && produced > crc_from) {
        z->crc = (uint32_t) crc32(z->crc, base + crc_from, (uInt) (produced - crc_from));
        z->member_bytes += produced - crc_from;
    }
    memcpy(dst, base, produced);
    /* history for the next piece */
    if (produced >= WINDOW) memcpy(z->out, base + produced - WINDOW, WINDOW);
    else if (produced > 0) {
        memmove(z->out, z->out + produced, WINDOW - produced);
        memcpy(z->out + WINDOW - produced, base, produced);
    }
    if (produced == 0 && !z->finished) return fail(z, "no progress"), -1;
    return (long) produced;
}

/* Test hook (include/hfg_io.h): the whole file through the decoder in pieces of piece_bytes. */
int hfg_debug_gunzip(const char *path, uint8_t *out, size_t cap, size_t *len, size_t piece_bytes, char *err, size_t errlen) {
    *len = 0;
    hfg_inflate *z = hfg_inflate_open(path, piece_bytes ? piece_bytes : (size_t) 4 << 20);
    if (!z) {
        snprintf(err, errlen, "%s: not a readable gzip file", path);
        return 1;
    }
    uint8_t *piece = malloc(hfg_inflate_piece_capacity(z));
    long got = piece ? 0 : -1;
    while (piece && (got = hfg_inflate_next(z, piece)) > 0) {
        if (*len + (size_t) got > cap) {
            got = -1;
            snprintf(z->err, sizeof(z->err), "output larger than the buffer");
            break;
        }
        memcpy(out + *len, piece, (size_t) got);
        *len += (size_t) got;
    }
    if (got < 0) snprintf(err, errlen, "%s: %s", path, piece ? hfg_inflate_error(z) : "out of memory");
    free(piece);
    hfg_inflate_close(z);
    return got < 0 ? 1 : 0;
}
