/*
 * jpgenc_oracle.c — CPU restatement (plain C) of the Nuos/jpgEnc encode path.
 * TEST INFRASTRUCTURE ONLY; see jpgenc_oracle.h for the rules and the parity status (PINNED).
 *
 * Written from the reference's behaviour, not from its text: planes are flat arrays, the bit
 * container is a byte buffer, the Huffman build tracks packages as DAG nodes.  What IS reproduced to
 * the last bit is the arithmetic order (double, no contraction: build with -ffp-contract=off) and the
 * two library-defined orders the file bytes depend on: libstdc++'s
 * unordered_map<int,int> iteration order and std::priority_queue's heap order.
 */
#include "jpgenc_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* constants                                                                                  */
/* ------------------------------------------------------------------------------------------ */
const uint8_t jo_qtable_luma[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t jo_qtable_chroma[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                      24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                      99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                      99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* natural index (row*8+col) of the i-th coefficient of the zigzag scan (Coding.hpp:57-81) */
static const uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                               12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                               58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

int jo_zigzag_index(int i) { return ZZ[i & 63]; }

/* AAN constants, evaluated with the reference's expressions (Dct.hpp:21-43) */
static double A1, A2, A3, A4, A5, S[8];
static int consts_ready = 0;
static void init_consts(void) {
    if (consts_ready) return;
    const double pi = 3.141592653589793238462643383279502884;
    const double root_two = 1.414213562373095048801688724209698078;
    double c[8];
    for (int k = 1; k < 8; ++k) c[k] = cos(k * pi / 16);
    A1 = c[4];
    A2 = c[2] - c[6];
    A3 = c[4];
    A4 = c[6] + c[2];
    A5 = c[6];
    S[0] = 1 / (2 * root_two);
    for (int k = 1; k < 8; ++k) S[k] = 1 / (4 * c[k]);
    consts_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* PPM                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const uint8_t* p; size_t pos, n; } reader;

static int rd_byte(reader* r) { return r->pos < r->n ? r->p[r->pos++] : (r->pos++, 0); }
static int rd_eof(const reader* r) { return r->pos >= r->n; }

/* PPMFileBuffer::read_word (Image.cpp:349-374): returns word bounds [*a, *b) */
static void rd_word(reader* r, size_t* a, size_t* b) {
    int c = rd_byte(r);
    if (isspace(c)) {
        while (r->pos < r->n && isspace(r->p[r->pos])) ++r->pos;
        c = rd_byte(r);
    }
    size_t first = r->pos - 1;
    for (;;) {
        if (c == '#') {
            while (!rd_eof(r) && r->p[r->pos++] != '\n') {}
            first = r->pos;
        } else if (isspace(c)) {
            *a = first; *b = r->pos - 1;
            return;
        } else if (rd_eof(r)) {
            *a = first; *b = r->pos < r->n ? r->pos : r->n;
            return;
        }
        c = rd_byte(r);
    }
}

static long word_to_int(const reader* r, size_t a, size_t b) {
    long v = 0;
    for (size_t i = a; i < b && i < r->n; ++i) v = v * 10 + (r->p[i] - '0');
    return v;
}

int jo_ppm_parse(const uint8_t* file, size_t n, jo_ppm_header* h) {
    reader r = {file, 0, n};
    size_t a, b;
    rd_word(&r, &a, &b);
    if (b - a != 2 || file[a] != 'P' || (file[a + 1] != '3' && file[a + 1] != '6')) return -1;
    h->magic = file[a + 1] - '0';
    rd_word(&r, &a, &b); h->width = (uint32_t)word_to_int(&r, a, b);
    rd_word(&r, &a, &b); h->height = (uint32_t)word_to_int(&r, a, b);
    rd_word(&r, &a, &b); h->maxval = (uint32_t)word_to_int(&r, a, b);
    h->payload = r.pos;
    if (r.pos > n) return -2;
    return 0;
}

int jo_ppm_samples(const uint8_t* file, size_t n, const jo_ppm_header* h, uint8_t* rgb) {
    const size_t count = (size_t)h->width * h->height * 3;
    if (h->magic == 6) {
        if (h->payload + count > n) return -2;
        memcpy(rgb, file + h->payload, count);
        return 0;
    }
    reader r = {file, h->payload, n};
    for (size_t i = 0; i < count; ++i) {
        size_t a, b;
        if (rd_eof(&r)) return -2;
        rd_word(&r, &a, &b);
        rgb[i] = (uint8_t)word_to_int(&r, a, b);   /* fast_atoi, Image.cpp:327-333 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* DCT variants                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* one 8-point AAN pass with the reference's operation order (Dct.hpp:52-131) */
static void aan8(const double x[8], double o[8]) {
    const double z0 = x[0] + x[7], z1 = x[1] + x[6], z2 = x[2] + x[5], z3 = x[3] + x[4];
    const double z4 = -x[4] + x[3], z5 = -x[5] + x[2], z6 = -x[6] + x[1], z7 = -x[7] + x[0];

    const double r0 = z0 + z3, r1 = z1 + z2, r2 = z1 - z2, r3 = z0 - z3;
    const double r4 = -z4 - z5, r5 = z5 + z6, r6 = z6 + z7, r7 = z7;

    const double t0 = r0 + r1, t1 = r0 - r1;
    double t2 = r2 + r3, t4 = r4, t5 = r5, t6 = r6;
    const double t3 = r3, t7 = r7;

    const double tmp = (t4 + t6) * A5;
    t2 *= A1; t4 *= A2; t5 *= A3; t6 *= A4;

    const double u4 = -t4 - tmp, u6 = t6 - tmp;
    const double v2 = t2 + t3, v3 = t3 - t2, v5 = t5 + t7, v7 = t7 - t5;
    const double w4 = u4 + v7, w5 = v5 + u6, w6 = -u6 + v5, w7 = v7 - u4;

    o[0] = t0 * S[0]; o[4] = t1 * S[4]; o[2] = v2 * S[2]; o[6] = v3 * S[6];
    o[5] = w4 * S[5]; o[1] = w5 * S[1]; o[7] = w6 * S[7]; o[3] = w7 * S[3];
}

void jo_dct_arai(const double in[64], double out[64]) {
    init_consts();
    double tmp[64], col[8];
    for (int j = 0; j < 8; ++j) {               /* pass 1: column j of the block -> row j of tmp */
        for (int i = 0; i < 8; ++i) col[i] = in[i * 8 + j];
        aan8(col, tmp + j * 8);
    }
    for (int j = 0; j < 8; ++j) {               /* pass 2: column j of tmp -> row j of the result */
        for (int i = 0; i < 8; ++i) col[i] = tmp[i * 8 + j];
        aan8(col, out + j * 8);
    }
}

static void dct_basis(double a[64]) {           /* Dct.hpp:219-235 */
    const double pi = 3.141592653589793238462643383279502884;
    const double root_two = 1.414213562373095048801688724209698078;
    const double scale = sqrt(2. / 8);
    for (unsigned k = 0; k < 8; ++k)
        for (unsigned n = 0; n < 8; ++n) {
            const double term = (2. * n + 1.) * ((k * pi) / 16.);
            a[k * 8 + n] = (k == 0 ? 1. / root_two : 1.) * scale * cos(term);
        }
}

void jo_dct_basis(double a[64]) { dct_basis(a); }

void jo_dct_direct(const double in[64], double out[64]) {   /* Dct.hpp:238-262 */
    double a[64];
    dct_basis(a);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double sum = 0.0;
            for (int x = 0; x < 8; ++x)
                for (int y = 0; y < 8; ++y) sum += in[y * 8 + x] * a[i * 8 + x] * a[j * 8 + y];
            out[j * 8 + i] = sum;
        }
}

void jo_dct_matrix(const double in[64], double out[64]) {   /* Y = A (X A^T), Dct.hpp:264-276 */
    double a[64], first[64];
    dct_basis(a);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += in[i * 8 + k] * a[j * 8 + k];
            first[i * 8 + j] = s;
        }
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double s = 0;
            for (int k = 0; k < 8; ++k) s += a[i * 8 + k] * first[k * 8 + j];
            out[i * 8 + j] = s;
        }
}

void jo_quantize(const double in[64], const uint8_t table[64], int32_t out[64]) {
    for (int i = 0; i < 64; ++i) out[i] = (int32_t)round(in[i] / (double)table[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* symbols                                                                                    */
/* ------------------------------------------------------------------------------------------ */
void jo_category(int value, int* category, uint32_t* bits) {
    if (value == 0) { *category = 0; *bits = 0; return; }
    const long a = labs((long)value);
    int cat = 1;
    while ((1L << cat) <= a) ++cat;             /* 2^(cat-1) <= |v| <= 2^cat - 1 */
    *category = cat;
    *bits = (uint32_t)(value < 0 ? ((1L << cat) - 1) - a : value);
}

/* zigzag-ordered block -> symbol list; zz[0] is whatever DC value should be coded */
/* okey[] (optional): order key of each entry inside its block, increasing along the text: DC 0, a ZRL 2p, the symbol of
 * the coefficient at zigzag position p 2p+1, EOB 129 (ZRLs of one run share a key; only their first one can matter) */
static int symbols_from_zigzag(const int32_t zz[64], uint8_t sym[64], uint32_t bits[64], uint8_t nbits[64], uint8_t* okey) {
    int n = 0, cat;
    uint32_t b;
    jo_category(zz[0], &cat, &b);
    if (okey) okey[n] = 0;
    sym[n] = (uint8_t)cat; bits[n] = b; nbits[n] = (uint8_t)cat; ++n;
    unsigned zeros = 0;
    for (int i = 1; i < 64; ++i) {
        if (zz[i] == 0) { ++zeros; continue; }
        while (zeros > 15) { if (okey) okey[n] = (uint8_t)(2 * i); sym[n] = 0xF0; bits[n] = 0; nbits[n] = 0; ++n; zeros -= 16; }
        jo_category(zz[i], &cat, &b);
        if (okey) okey[n] = (uint8_t)(2 * i + 1);
        sym[n] = (uint8_t)((zeros << 4) | cat); bits[n] = b; nbits[n] = (uint8_t)cat; ++n;
        zeros = 0;
    }
    if (zeros > 0) { if (okey) okey[n] = 129; sym[n] = 0; bits[n] = 0; nbits[n] = 0; ++n; }   /* EOB */
    return n;
}

int jo_block_symbols(const int32_t natural[64], uint8_t sym[64], uint32_t bits[64], uint8_t nbits[64]) {
    int32_t zz[64];
    for (int i = 0; i < 64; ++i) zz[i] = natural[ZZ[i]];
    return symbols_from_zigzag(zz, sym, bits, nbits, NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* libstdc++ unordered_map<int,int> iteration order (bits/hashtable.h, hashtable_policy.h)      */
/* ------------------------------------------------------------------------------------------ */
#define UM_MAXN 256
#define UM_MAXB 600
#define UM_BEFORE_BEGIN (-2)
typedef struct {
    int nb;                 /* bucket count */
    int n;                  /* element count */
    int next_resize;
    int head;               /* first node, -1 if empty */
    int key[UM_MAXN];
    int nxt[UM_MAXN];
    int slot_of[256];       /* key -> node index, -1 */
    int bucket[UM_MAXB];    /* node BEFORE the bucket's first node; -1 = empty bucket */
} umap;

static void um_init(umap* m) {
    m->nb = 1; m->n = 0; m->next_resize = 0; m->head = -1;
    memset(m->slot_of, 0xff, sizeof m->slot_of);
    m->bucket[0] = -1;
}

static int um_next_bkt(umap* m, int want) {     /* _Prime_rehash_policy::_M_next_bkt */
    static const unsigned char fast[] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13};
    static const int primes[] = {17,  19,  23,  29,  31,  37,  41,  43,  47,  53,  59,  61,  67,  71,
                                 73,  79,  83,  89,  97,  103, 109, 113, 127, 137, 139, 149, 157, 167,
                                 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409,
                                 439, 467, 503, 541};
    if (want < (int)sizeof fast) {
        if (want == 0) return 1;
        m->next_resize = fast[want];
        return fast[want];
    }
    for (size_t i = 0; i < sizeof primes / sizeof primes[0]; ++i)
        if (primes[i] >= want) { m->next_resize = primes[i]; return primes[i]; }
    abort();
}

static int um_get_next(const umap* m, int node) { return node == UM_BEFORE_BEGIN ? m->head : m->nxt[node]; }
static void um_set_next(umap* m, int node, int v) { if (node == UM_BEFORE_BEGIN) m->head = v; else m->nxt[node] = v; }

static void um_rehash(umap* m, int nb) {        /* _M_rehash_aux(n, true_type) */
    int p = m->head, bbegin_bkt = 0;
    for (int i = 0; i < nb; ++i) m->bucket[i] = -1;
    m->head = -1;
    while (p >= 0) {
        const int next = m->nxt[p];
        const int bkt = m->key[p] % nb;
        if (m->bucket[bkt] == -1) {
            m->nxt[p] = m->head;
            m->head = p;
            m->bucket[bkt] = UM_BEFORE_BEGIN;
            if (m->nxt[p] >= 0) m->bucket[bbegin_bkt] = p;
            bbegin_bkt = bkt;
        } else {
            const int before = m->bucket[bkt];
            m->nxt[p] = um_get_next(m, before);
            um_set_next(m, before, p);
        }
        p = next;
    }
    m->nb = nb;
}

/* returns the node of `key`, inserting it when new (operator[]) */
static int um_touch(umap* m, int key) {
    if (m->slot_of[key] >= 0) return m->slot_of[key];
    /* _M_need_rehash(n_bkt, n_elt, 1) */
    if (m->n + 1 > m->next_resize) {
        int min_bkts = m->n + 1;
        if (m->next_resize == 0 && min_bkts < 11) min_bkts = 11;
        if (min_bkts >= m->nb) {
            int want = min_bkts + 1;
            if (want < m->nb * 2) want = m->nb * 2;
            um_rehash(m, um_next_bkt(m, want));
        } else {
            m->next_resize = m->nb;
        }
    }
    const int node = m->n++;
    m->key[node] = key;
    m->slot_of[key] = node;
    const int bkt = key % m->nb;
    if (m->bucket[bkt] != -1) {                 /* _M_insert_bucket_begin */
        const int before = m->bucket[bkt];
        m->nxt[node] = um_get_next(m, before);
        um_set_next(m, before, node);
    } else {
        m->nxt[node] = m->head;
        m->head = node;
        if (m->nxt[node] >= 0) m->bucket[m->key[m->nxt[node]] % m->nb] = node;
        m->bucket[bkt] = UM_BEFORE_BEGIN;
    }
    return node;
}

/* ------------------------------------------------------------------------------------------ */
/* std::priority_queue order (bits/stl_heap.h), comparator = "weight greater" (Huffman.hpp:117)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int64_t w; int node; } hitem;
typedef struct { hitem* v; int n; } heap;

static void heap_push_hole(hitem* v, int hole, int top, hitem val) {    /* std::__push_heap */
    int parent = (hole - 1) / 2;
    while (hole > top && v[parent].w > val.w) {
        v[hole] = v[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[hole] = val;
}
static void heap_push(heap* h, hitem x) { h->v[h->n++] = x; heap_push_hole(h->v, h->n - 1, 0, x); }
static hitem heap_pop(heap* h) {                                       /* top() then pop() */
    hitem top = h->v[0];
    if (h->n > 1) {
        const int len = h->n - 1;                /* std::__pop_heap + std::__adjust_heap */
        hitem val = h->v[len];
        h->v[len] = h->v[0];
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (h->v[child].w > h->v[child - 1].w) --child;
            h->v[hole] = h->v[child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h->v[hole] = h->v[child - 1];
            hole = child - 1;
        }
        heap_push_hole(h->v, hole, 0, val);
    }
    --h->n;
    return top;
}

/* ------------------------------------------------------------------------------------------ */
/* Huffman table build                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int left, right, sym; } pnode;  /* leaf: sym>=0; package: children */

static void count_leaves(const pnode* nodes, int root, int cnt[256], uint8_t present[256]) {
    int stack[64 * 16], sp = 0;                   /* depth <= 16; explicit stack of pending nodes */
    int* big = NULL;
    int cap = (int)(sizeof stack / sizeof stack[0]);
    int* st = stack;
    st[sp++] = root;
    while (sp) {
        const int k = st[--sp];
        if (nodes[k].sym >= 0) { ++cnt[nodes[k].sym]; present[nodes[k].sym] = 1; continue; }
        if (sp + 2 > cap) {
            cap *= 2;
            int* nb = (int*)malloc(sizeof(int) * cap);
            memcpy(nb, st, sizeof(int) * sp);
            free(big);
            big = nb; st = nb;
        }
        st[sp++] = nodes[k].left;
        st[sp++] = nodes[k].right;
    }
    free(big);
}

/* symbols/frequencies in unordered_map iteration order -> table (package_merge, limit 15; Huffman.hpp:114-174) */
static void build_from_ordered(const int* syms, const uint32_t* freq, int n, jo_huff_table* t) {
    memset(t, 0, sizeof *t);
    t->nsymbols = n;
    int per_len[18][256], per_len_n[18];
    memset(per_len_n, 0, sizeof per_len_n);

    if (n == 1) {                                /* Huffman.cpp:17-25 */
        per_len[1][per_len_n[1]++] = syms[0];
    } else {
        const int limit = 15;
        const int max_nodes = n * (limit + 2) * 2 + 16;
        pnode* nodes = (pnode*)malloc(sizeof(pnode) * max_nodes);
        int nn = 0;
        heap blueprint = {(hitem*)malloc(sizeof(hitem) * n), 0};
        for (int i = 0; i < n; ++i) {
            nodes[nn].left = nodes[nn].right = -1; nodes[nn].sym = syms[i];
            hitem it = {(int64_t)freq[i], nn++};
            heap_push(&blueprint, it);
        }
        heap cur = {(hitem*)malloc(sizeof(hitem) * (2 * n + 16)), 0};
        heap nxt = {(hitem*)malloc(sizeof(hitem) * (2 * n + 16)), 0};
        memcpy(cur.v, blueprint.v, sizeof(hitem) * n); cur.n = n;
        for (int lvl = 0; lvl < limit; ++lvl) {
            if (lvl + 1 < limit) { memcpy(nxt.v, blueprint.v, sizeof(hitem) * n); nxt.n = n; }
            else nxt.n = 0;                       /* last level starts empty (Huffman.hpp:133) */
            while (cur.n > 1) {
                hitem p1 = heap_pop(&cur);
                hitem p2 = heap_pop(&cur);
                nodes[nn].left = p1.node; nodes[nn].right = p2.node; nodes[nn].sym = -1;
                hitem m = {p1.w + p2.w, nn++};
                heap_push(&nxt, m);
            }
            heap sw = cur; cur = nxt; nxt = sw;
        }
        /* cur == final level: pop packages, count symbol occurrences, remember first-touch order */
        umap lengths_map;
        um_init(&lengths_map);
        int len_of[256];
        memset(len_of, 0, sizeof len_of);
        while (cur.n) {
            hitem p = heap_pop(&cur);
            int cnt[256];
            uint8_t present[256];
            memset(cnt, 0, sizeof cnt);
            memset(present, 0, sizeof present);
            count_leaves(nodes, p.node, cnt, present);
            for (int s = 0; s < 256; ++s)         /* a package lists its symbols in ascending order */
                if (present[s]) { um_touch(&lengths_map, s); len_of[s] += cnt[s]; }
        }
        for (int p = lengths_map.head; p >= 0; p = lengths_map.nxt[p]) {
            const int s = lengths_map.key[p];
            per_len[len_of[s]][per_len_n[len_of[s]]++] = s;
        }
        /* preventOnlyOnesCode (Huffman.cpp:37-48) */
        int deepest = 16;
        while (deepest > 0 && per_len_n[deepest] == 0) --deepest;
        const int moved = per_len[deepest][--per_len_n[deepest]];
        per_len[deepest + 1][per_len_n[deepest + 1]++] = moved;
        free(nodes); free(blueprint.v); free(cur.v); free(nxt.v);
    }
    /* generateCodes (Huffman.cpp:50-66) */
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        t->counts[len - 1] = (uint8_t)per_len_n[len];
        for (int i = 0; i < per_len_n[len]; ++i) {
            const int s = per_len[len][i];
            t->code_msb[s] = code << (32 - len);
            t->length[s] = (uint8_t)len;
            t->symbols[k++] = (uint8_t)s;
            ++code;
        }
        code <<= 1;
    }
}

void jo_huffman_from_hist(const uint32_t count[256], const uint8_t* order, int ndistinct, jo_huff_table* t) {
    umap m;
    um_init(&m);
    for (int i = 0; i < ndistinct; ++i) um_touch(&m, order[i]);
    int syms[256], n = 0;
    uint32_t freq[256];
    for (int p = m.head; p >= 0; p = m.nxt[p]) { syms[n] = m.key[p]; freq[n] = count[m.key[p]]; ++n; }
    build_from_ordered(syms, freq, n, t);
}

void jo_huffman_from_text(const int32_t* text, size_t n, jo_huff_table* t) {
    uint32_t count[256];
    uint8_t order[256];
    int nd = 0;
    memset(count, 0, sizeof count);
    for (size_t i = 0; i < n; ++i) {
        const int s = text[i] & 255;
        if (count[s]++ == 0) order[nd++] = (uint8_t)s;
    }
    jo_huffman_from_hist(count, order, nd, t);
}

/* ------------------------------------------------------------------------------------------ */
/* bit container                                                                              */
/* ------------------------------------------------------------------------------------------ */
void jo_bits_init(jo_bits* b) { b->bytes = NULL; b->cap = 0; b->nbits = 0; }
void jo_bits_free(jo_bits* b) { free(b->bytes); jo_bits_init(b); }

static void bits_reserve(jo_bits* b, uint64_t more_bits) {
    const size_t need = (size_t)((b->nbits + more_bits + 7) / 8) + 8;
    if (need <= b->cap) return;
    size_t cap = b->cap ? b->cap : 256;
    while (cap < need) cap *= 2;
    b->bytes = (uint8_t*)realloc(b->bytes, cap);
    memset(b->bytes + b->cap, 0, cap - b->cap);
    b->cap = cap;
}

/* append the low n bits of value, most significant of them first (n <= 32) */
void jo_bits_push_lsb(jo_bits* b, uint32_t value, int n) {
    if (n <= 0) return;
    bits_reserve(b, (uint64_t)n);
    uint64_t pos = b->nbits;
    for (int i = n - 1; i >= 0; --i, ++pos)
        if ((value >> i) & 1u) b->bytes[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
    b->nbits = pos;
}

void jo_bits_push_msb(jo_bits* b, uint32_t msb_aligned, int n) {
    if (n <= 0) return;
    jo_bits_push_lsb(b, n >= 32 ? msb_aligned : msb_aligned >> (32 - n), n);
}

void jo_bits_fill(jo_bits* b) {
    /* fill() pads the open byte with 1s; on an EMPTY stream bit_idx==8 makes it emit a whole 0xFF
     * byte (BitstreamGeneric.hpp:242-248 with the ctor's bit_idx = block_size) */
    if (b->nbits == 0) { jo_bits_push_lsb(b, 0xFF, 8); return; }
    const int rem = (int)(b->nbits & 7);
    if (rem) jo_bits_push_lsb(b, 0xFFu, 8 - rem);
}

size_t jo_bits_stuffed_size(const jo_bits* b) {
    const size_t nbytes = (size_t)((b->nbits + 7) / 8);
    size_t extra = 0;
    for (size_t i = 0; i < nbytes; ++i) extra += b->bytes[i] == 0xFF;
    return nbytes + extra;
}

size_t jo_bits_write_stuffed(const jo_bits* b, uint8_t* dst) {
    const size_t nbytes = (size_t)((b->nbits + 7) / 8);
    size_t o = 0;
    for (size_t i = 0; i < nbytes; ++i) {
        dst[o++] = b->bytes[i];
        if (b->bytes[i] == 0xFF) dst[o++] = 0x00;
    }
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* forward path                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* colour conversion of one (already scaled) pixel, Image.cpp:131-143: float constants, double maths */
static inline void rgb_to_ycc(double r, double g, double b, double* y, double* cb, double* cr) {
    static const float Flat[3] = {.0f, 256 / 2.f, 256 / 2.f};
    static const float Yv[3] = {.299f, .587f, .114f};
    static const float Cbv[3] = {-.1687f, -.3312f, .5f};
    static const float Crv[3] = {.5f, -.4186f, -.0813f};
    *y  = Flat[0] + (Yv[0] * r + Yv[1] * g + Yv[2] * b) - 128;
    *cb = Flat[1] + (Cbv[0] * r + Cbv[1] * g + Cbv[2] * b) - 128;
    *cr = Flat[2] + (Crv[0] * r + Crv[1] * g + Crv[2] * b) - 128;
}

/* converts MCU row `my` into y[16][W16], cb[8][W16/2], cr[8][W16/2] */
static void convert_mcu_row(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, double scale,
                            uint32_t W16, uint32_t my, double* y, double* cbf, double* crf,
                            double* cb, double* cr) {
    for (uint32_t r = 0; r < 16; ++r) {
        uint32_t sy = my * 16 + r;
        if (sy >= real_h) sy = real_h - 1;                     /* edge replication, Image.cpp:491-530 */
        const uint8_t* row = rgb + (size_t)sy * real_w * 3;
        for (uint32_t x = 0; x < W16; ++x) {
            const uint32_t sx = x < real_w ? x : real_w - 1;
            const double R = row[sx * 3 + 0] * scale, G = row[sx * 3 + 1] * scale, B = row[sx * 3 + 2] * scale;
            rgb_to_ycc(R, G, B, &y[r * W16 + x], &cbf[r * W16 + x], &crf[r * W16 + x]);
        }
    }
    const uint32_t W8 = W16 / 2;
    for (uint32_t r = 0; r < 8; ++r)                            /* S420_m, Image.cpp:207-226 */
        for (uint32_t x = 0; x < W8; ++x) {
            const double* a = cbf + (2 * r) * W16 + 2 * x;
            double top = 0; top += 1 * a[0]; top += 1 * a[1];
            double bot = 0; bot += 1 * a[W16]; bot += 1 * a[W16 + 1];
            cb[r * W8 + x] = (top + bot) / 4;
            a = crf + (2 * r) * W16 + 2 * x;
            top = 0; top += 1 * a[0]; top += 1 * a[1];
            bot = 0; bot += 1 * a[W16]; bot += 1 * a[W16 + 1];
            cr[r * W8 + x] = (top + bot) / 4;
        }
}

static void block_from(const double* plane, uint32_t pitch, uint32_t x0, uint32_t y0, double blk[64]) {
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) blk[i * 8 + j] = plane[(size_t)(y0 + i) * pitch + x0 + j];
}

static void emit_block(const double blk[64], const uint8_t q[64], int16_t* out, double* dct_dump, int32_t* q_dump,
                       uint32_t pitch) {
    double d[64];
    int32_t qi[64];
    jo_dct_arai(blk, d);
    jo_quantize(d, q, qi);
    if (out) for (int i = 0; i < 64; ++i) out[i] = (int16_t)qi[ZZ[i]];
    if (dct_dump) for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) dct_dump[(size_t)i * pitch + j] = d[i * 8 + j];
    if (q_dump) for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) q_dump[(size_t)i * pitch + j] = qi[i * 8 + j];
}

static void forward_impl(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval,
                         const uint8_t qy[64], const uint8_t qc[64], uint32_t my0, uint32_t my1, int16_t* out,
                         double* py, double* pcb, double* pcr, double* dy, double* dcb, double* dcr,
                         int32_t* qyp, int32_t* qcbp, int32_t* qcrp) {
    init_consts();
    const uint32_t W16 = jo_pad16(real_w), W8 = W16 / 2, mcu_w = W16 / 16;
    const double scale = 255. / maxval;                         /* Image.cpp:465 */
    double* y = (double*)malloc(sizeof(double) * 16 * W16);
    double* cbf = (double*)malloc(sizeof(double) * 16 * W16);
    double* crf = (double*)malloc(sizeof(double) * 16 * W16);
    double* cb = (double*)malloc(sizeof(double) * 8 * W8);
    double* cr = (double*)malloc(sizeof(double) * 8 * W8);
    double blk[64];
    for (uint32_t my = my0; my < my1; ++my) {
        convert_mcu_row(rgb, real_w, real_h, scale, W16, my, y, cbf, crf, cb, cr);
        if (py) memcpy(py + (size_t)my * 16 * W16, y, sizeof(double) * 16 * W16);
        if (pcb) memcpy(pcb + (size_t)my * 8 * W8, cb, sizeof(double) * 8 * W8);
        if (pcr) memcpy(pcr + (size_t)my * 8 * W8, cr, sizeof(double) * 8 * W8);
        for (uint32_t mx = 0; mx < mcu_w; ++mx) {
            int16_t* o = out ? out + ((size_t)(my - my0) * mcu_w + mx) * 6 * 64 : NULL;
            for (int k = 0; k < 4; ++k) {
                const uint32_t bx = mx * 16 + (k & 1) * 8, by = (k >> 1) * 8;
                block_from(y, W16, bx, by, blk);
                const size_t off = ((size_t)my * 16 + by) * W16 + bx;
                emit_block(blk, qy, o ? o + k * 64 : NULL, dy ? dy + off : NULL, qyp ? qyp + off : NULL, W16);
            }
            const size_t coff = (size_t)my * 8 * W8 + mx * 8;
            block_from(cb, W8, mx * 8, 0, blk);
            emit_block(blk, qc, o ? o + 4 * 64 : NULL, dcb ? dcb + coff : NULL, qcbp ? qcbp + coff : NULL, W8);
            block_from(cr, W8, mx * 8, 0, blk);
            emit_block(blk, qc, o ? o + 5 * 64 : NULL, dcr ? dcr + coff : NULL, qcrp ? qcrp + coff : NULL, W8);
        }
    }
    free(y); free(cbf); free(crf); free(cb); free(cr);
}

void jo_forward_mcu_rows(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval,
                         const uint8_t qy[64], const uint8_t qc[64], uint32_t mcu_y0, uint32_t mcu_y1, int16_t* out) {
    forward_impl(rgb, real_w, real_h, maxval, qy, qc, mcu_y0, mcu_y1, out, 0, 0, 0, 0, 0, 0, 0, 0, 0);
}

void jo_forward_planes(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval,
                       const uint8_t qy[64], const uint8_t qc[64], double* y, double* cb, double* cr,
                       double* dct_y, double* dct_cb, double* dct_cr, int32_t* q_y, int32_t* q_cb, int32_t* q_cr) {
    forward_impl(rgb, real_w, real_h, maxval, qy, qc, 0, jo_pad16(real_h) / 16, NULL, y, cb, cr, dct_y, dct_cb,
                 dct_cr, q_y, q_cb, q_cr);
}

/* Image::subsample for every SubsamplingMode (Image.cpp:198-235, masks and divisors :237-319): mode 0 S444, 1 S422, 2 S411,
 * 3 S420, 4 S420_m, 5 S420_lm (the enum order of Image.hpp).  in: h x w doubles; out: (h / vdiv) x (w / hdiv).  The weights
 * are applied the way the reference applies them -- pix = 0; pix += row[m] * chan(y, x + m) for EVERY tap, zero weights
 * included -- and the second scanline is added and divided in place. */
void jo_subsample_dims(int mode, uint32_t w, uint32_t h, uint32_t* ow, uint32_t* oh) {
    const int hdiv = (mode == 0) ? 1 : (mode == 2) ? 4 : 2;
    const int vdiv = (mode >= 3) ? 2 : 1;
    *ow = w / hdiv; *oh = h / vdiv;
}

void jo_subsample_plane(const double* in, uint32_t w, uint32_t h, int mode, double* out) {
    if (mode == 0) { memcpy(out, in, sizeof(double) * (size_t)w * h); return; }
    const int taps = (mode == 2) ? 4 : 2;
    const double row[4] = {1, (mode == 4) ? 1 : 0, 0, 0};      /* {1,0} {1,0,0,0} {1,0} {1,1} {1,0} */
    const int scanline_jump = (mode == 3), averaging = (mode == 4 || mode == 5);
    const double div = (mode == 4) ? 4 : 2;
    size_t pixidx = 0, pixidx2 = 0;
    for (uint32_t y = 0; y < h; y += 2) {
        for (uint32_t x = 0; x < w; x += taps) {
            double pix = 0;
            for (int m = 0; m < taps; ++m) pix += row[m] * in[(size_t)y * w + x + m];
            out[pixidx++] = pix;
        }
        if (!scanline_jump) {
            if (averaging) {
                for (uint32_t x = 0; x < w; x += taps) {
                    double pix = 0;
                    for (int m = 0; m < taps; ++m) pix += row[m] * in[(size_t)(y + 1) * w + x + m];
                    out[pixidx2] += pix;
                    out[pixidx2] /= div;
                    ++pixidx2;
                }
            } else {
                --y;                                            /* the next scanline is an output row of its own */
            }
        }
    }
}

/* Image::applyDCT on one plane (Image.cpp:540-595): every 8x8 block through dctDirect / dctMat / dctArai (mode 0 / 1 / 2) */
void jo_dct_plane(const double* in, uint32_t w, uint32_t h, int mode, double* out) {
    double a[64], b[64];
    for (uint32_t by = 0; by < h; by += 8)
        for (uint32_t bx = 0; bx < w; bx += 8) {
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) a[i * 8 + j] = in[(size_t)(by + i) * w + bx + j];
            if (mode == 0) jo_dct_direct(a, b); else if (mode == 1) jo_dct_matrix(a, b); else jo_dct_arai(a, b);
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) out[(size_t)(by + i) * w + bx + j] = b[i * 8 + j];
        }
}

/* Image::writeJPEG on an Image whose planes are given as doubles (Image.hpp:104-114): p0,p1,p2 are H16*W16 row-major planes,
 * R,G,B -- converted like any pixel (Image.cpp:131-143) -- or, with ycbcr != 0, level-shifted Y,Cb,Cr that
 * convertToColorSpace leaves alone (Image.cpp:112-115); then S420_m (Image.cpp:198-235), dctArai, quantize as above. */
void jo_forward_from_planes(const double* p0, const double* p1, const double* p2, uint32_t W16, uint32_t H16, int ycbcr,
                            const uint8_t qy[64], const uint8_t qc[64], int16_t* out) {
    init_consts();
    const uint32_t W8 = W16 / 2, mcu_w = W16 / 16, mcu_h = H16 / 16;
    double* y = (double*)malloc(sizeof(double) * 16 * W16);
    double* cbf = (double*)malloc(sizeof(double) * 16 * W16);
    double* crf = (double*)malloc(sizeof(double) * 16 * W16);
    double* cb = (double*)malloc(sizeof(double) * 8 * W8);
    double* cr = (double*)malloc(sizeof(double) * 8 * W8);
    double blk[64];
    for (uint32_t my = 0; my < mcu_h; ++my) {
        for (uint32_t r = 0; r < 16; ++r)
            for (uint32_t x = 0; x < W16; ++x) {
                const size_t i = ((size_t)my * 16 + r) * W16 + x;
                if (ycbcr) { y[r * W16 + x] = p0[i]; cbf[r * W16 + x] = p1[i]; crf[r * W16 + x] = p2[i]; }
                else rgb_to_ycc(p0[i], p1[i], p2[i], &y[r * W16 + x], &cbf[r * W16 + x], &crf[r * W16 + x]);
            }
        for (uint32_t r = 0; r < 8; ++r)                            /* S420_m, Image.cpp:207-226 */
            for (uint32_t x = 0; x < W8; ++x) {
                const double* a = cbf + (2 * r) * W16 + 2 * x;
                double top = 0; top += 1 * a[0]; top += 1 * a[1];
                double bot = 0; bot += 1 * a[W16]; bot += 1 * a[W16 + 1];
                cb[r * W8 + x] = (top + bot) / 4;
                a = crf + (2 * r) * W16 + 2 * x;
                top = 0; top += 1 * a[0]; top += 1 * a[1];
                bot = 0; bot += 1 * a[W16]; bot += 1 * a[W16 + 1];
                cr[r * W8 + x] = (top + bot) / 4;
            }
        for (uint32_t mx = 0; mx < mcu_w; ++mx) {
            int16_t* o = out + ((size_t)my * mcu_w + mx) * 6 * 64;
            for (int k = 0; k < 4; ++k) {
                block_from(y, W16, mx * 16 + (k & 1) * 8, (k >> 1) * 8, blk);
                emit_block(blk, qy, o + k * 64, NULL, NULL, W16);
            }
            block_from(cb, W8, mx * 8, 0, blk);
            emit_block(blk, qc, o + 4 * 64, NULL, NULL, W8);
            block_from(cr, W8, mx * 8, 0, blk);
            emit_block(blk, qc, o + 5 * 64, NULL, NULL, W8);
        }
    }
    free(y); free(cbf); free(crf); free(cb); free(cr);
}

void jo_planes_to_mcu(const int32_t* q_y, const int32_t* q_cb, const int32_t* q_cr, uint32_t mcu_w, uint32_t mcu_h,
                      int16_t* out) {
    const uint32_t W16 = mcu_w * 16, W8 = mcu_w * 8;
    for (uint32_t my = 0; my < mcu_h; ++my)
        for (uint32_t mx = 0; mx < mcu_w; ++mx) {
            int16_t* o = out + ((size_t)my * mcu_w + mx) * 6 * 64;
            for (int k = 0; k < 6; ++k) {
                const int32_t* src;
                uint32_t pitch;
                if (k < 4) { src = q_y + ((size_t)my * 16 + (k >> 1) * 8) * W16 + mx * 16 + (k & 1) * 8; pitch = W16; }
                else { src = (k == 4 ? q_cb : q_cr) + (size_t)my * 8 * W8 + mx * 8; pitch = W8; }
                for (int i = 0; i < 64; ++i) o[k * 64 + i] = (int16_t)src[(size_t)(ZZ[i] >> 3) * pitch + (ZZ[i] & 7)];
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* entropy coding                                                                             */
/* ------------------------------------------------------------------------------------------ */
/* per-block symbol list with the DC replaced by its difference.  prev[] = running predictors:
 * [0] Y in MCU order (Image.cpp:640-659), [1] Cb, [2] Cr raster (Image.cpp:661-677) */
static int mcu_block_symbols(const int16_t* blk, int comp, int32_t prev[3], uint8_t sym[64], uint32_t bits[64],
                             uint8_t nbits[64], uint8_t* okey) {
    int32_t zz[64];
    for (int i = 0; i < 64; ++i) zz[i] = blk[i];
    const int32_t dc = zz[0];
    zz[0] = dc - prev[comp];
    prev[comp] = dc;
    return symbols_from_zigzag(zz, sym, bits, nbits, okey);
}

void jo_symbol_stats(const int16_t* mcu_blocks, uint32_t mcu_w, uint32_t mcu_h, uint32_t count[4][256],
                     uint64_t first_pos[4][256]) {
    memset(count, 0, sizeof(uint32_t) * 4 * 256);
    memset(first_pos, 0xff, sizeof(uint64_t) * 4 * 256);
    int32_t prev[3] = {0, 0, 0};
    uint8_t sym[64], nb[64], okey[64];
    uint32_t bits[64];
    const uint64_t ncb = (uint64_t)mcu_w * mcu_h;
    for (uint32_t my = 0; my < mcu_h; ++my)
        for (uint32_t mx = 0; mx < mcu_w; ++mx) {
            const int16_t* m = mcu_blocks + ((size_t)my * mcu_w + mx) * 6 * 64;
            for (int k = 0; k < 6; ++k) {
                const int comp = k < 4 ? 0 : k - 3;
                const int n = mcu_block_symbols(m + k * 64, comp, prev, sym, bits, nb, okey);
                /* text order (Image.cpp:892-906): Y blocks raster over the block grid; chroma = all Cb then all Cr */
                uint64_t blk_index;
                if (k < 4) blk_index = ((uint64_t)my * 2 + (k >> 1)) * (mcu_w * 2) + mx * 2 + (k & 1);
                else blk_index = (uint64_t)(k - 4) * ncb + (uint64_t)my * mcu_w + mx;
                const int tdc = k < 4 ? 0 : 2, tac = tdc + 1;
                ++count[tdc][sym[0]];
                if (blk_index * 256 < first_pos[tdc][sym[0]]) first_pos[tdc][sym[0]] = blk_index * 256;
                for (int i = 1; i < n; ++i) {
                    ++count[tac][sym[i]];
                    const uint64_t key = blk_index * 256 + okey[i];
                    if (key < first_pos[tac][sym[i]]) first_pos[tac][sym[i]] = key;
                }
            }
        }
}

static void table_from_stats(const uint32_t count[256], const uint64_t first_pos[256], jo_huff_table* t) {
    uint8_t order[256];
    int nd = 0;
    for (int s = 0; s < 256; ++s) if (count[s]) order[nd++] = (uint8_t)s;
    for (int i = 1; i < nd; ++i) {                /* insertion sort by first occurrence */
        const uint8_t s = order[i];
        int j = i;
        while (j > 0 && first_pos[order[j - 1]] > first_pos[s]) { order[j] = order[j - 1]; --j; }
        order[j] = s;
    }
    jo_huffman_from_hist(count, order, nd, t);
}

void jo_entropy_encode(const int16_t* mcu_blocks, uint32_t mcu_w, uint32_t mcu_h, jo_huff_table tables[4],
                       jo_bits* out) {
    uint32_t count[4][256];
    uint64_t first_pos[4][256];
    jo_symbol_stats(mcu_blocks, mcu_w, mcu_h, count, first_pos);
    for (int t = 0; t < 4; ++t) table_from_stats(count[t], first_pos[t], &tables[t]);

    int32_t prev[3] = {0, 0, 0};
    uint8_t sym[64], nb[64];
    uint32_t bits[64];
    const size_t nmcu = (size_t)mcu_w * mcu_h;
    for (size_t m = 0; m < nmcu; ++m)             /* Image.cpp:959-967 */
        for (int k = 0; k < 6; ++k) {
            const int comp = k < 4 ? 0 : k - 3;
            const int n = mcu_block_symbols(mcu_blocks + (m * 6 + k) * 64, comp, prev, sym, bits, nb, NULL);
            const jo_huff_table* dc = &tables[k < 4 ? 0 : 2];
            const jo_huff_table* ac = &tables[k < 4 ? 1 : 3];
            jo_bits_push_msb(out, dc->code_msb[sym[0]], dc->length[sym[0]]);   /* Image.cpp:757-759 */
            jo_bits_push_lsb(out, bits[0], nb[0]);
            for (int i = 1; i < n; ++i) {
                jo_bits_push_msb(out, ac->code_msb[sym[i]], ac->length[sym[i]]);
                jo_bits_push_lsb(out, bits[i], nb[i]);
            }
        }
    jo_bits_fill(out);
}

/* ------------------------------------------------------------------------------------------ */
/* file assembly                                                                              */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint8_t* p; size_t n; } wr;
static void w8(wr* w, unsigned v) { if (w->p) w->p[w->n] = (uint8_t)v; ++w->n; }
static void w16(wr* w, unsigned v) { w8(w, (v >> 8) & 0xFF); w8(w, v & 0xFF); }   /* getHi/getLo on a short */

size_t jo_write_headers(uint32_t real_w, uint32_t real_h, const uint8_t qy[64], const uint8_t qc[64],
                        const jo_huff_table tables[4], uint8_t* dst) {
    wr w = {dst, 0};
    w16(&w, 0xFFD8);                                                     /* sSOI */
    w16(&w, 0xFFE0); w16(&w, 16);                                         /* sAPP0, JpegSegments.hpp:73-108 */
    w8(&w, 'J'); w8(&w, 'F'); w8(&w, 'I'); w8(&w, 'F'); w8(&w, 0);
    w8(&w, 1); w8(&w, 1); w8(&w, 0); w16(&w, 1); w16(&w, 1); w8(&w, 0); w8(&w, 0);
    for (int t = 0; t < 2; ++t) {                                         /* sDQT x2, one table each */
        const uint8_t* q = t ? qc : qy;
        w16(&w, 0xFFDB); w16(&w, 2 + 65); w8(&w, t);
        for (int i = 0; i < 64; ++i) w8(&w, q[ZZ[i]]);
    }
    w16(&w, 0xFFC0); w16(&w, 8 + 3 * 3); w8(&w, 8);                        /* sSOF0: Y then X */
    w16(&w, real_h & 0xFFFF); w16(&w, real_w & 0xFFFF); w8(&w, 3);
    w8(&w, 1); w8(&w, 0x22); w8(&w, 0);
    w8(&w, 2); w8(&w, 0x11); w8(&w, 1);
    w8(&w, 3); w8(&w, 0x11); w8(&w, 1);
    static const uint8_t info[4] = {0x00, 0x10, 0x01, 0x11};              /* (class<<4)|dest, Image.cpp:946-949 */
    for (int t = 0; t < 4; ++t) {
        const jo_huff_table* h = &tables[t];
        int nsym = 0;
        for (int i = 0; i < 16; ++i) nsym += h->counts[i];
        w16(&w, 0xFFC4); w16(&w, 2 + 17 + nsym); w8(&w, info[t]);
        for (int i = 0; i < 16; ++i) w8(&w, h->counts[i]);
        for (int i = 0; i < nsym; ++i) w8(&w, h->symbols[i]);
    }
    w16(&w, 0xFFDA); w16(&w, 12); w8(&w, 3);                              /* sSOS */
    w8(&w, 1); w8(&w, 0x00); w8(&w, 2); w8(&w, 0x11); w8(&w, 3); w8(&w, 0x11);
    w8(&w, 0x00); w8(&w, 0x3F); w8(&w, 0x00);
    return w.n;
}

void jo_free(void* p) { free(p); }

static int encode_common(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint32_t maxval, uint8_t** out,
                         size_t* out_n) {
    const uint32_t mcu_w = jo_pad16(real_w) / 16, mcu_h = jo_pad16(real_h) / 16;
    int16_t* coef = (int16_t*)malloc(sizeof(int16_t) * (size_t)mcu_w * mcu_h * 6 * 64);
    jo_forward_mcu_rows(rgb, real_w, real_h, maxval, jo_qtable_luma, jo_qtable_chroma, 0, mcu_h, coef);
    jo_huff_table tables[4];
    jo_bits bits;
    jo_bits_init(&bits);
    jo_entropy_encode(coef, mcu_w, mcu_h, tables, &bits);
    free(coef);
    const size_t hdr = jo_write_headers(real_w, real_h, jo_qtable_luma, jo_qtable_chroma, tables, NULL);
    const size_t scan = jo_bits_stuffed_size(&bits);
    uint8_t* buf = (uint8_t*)malloc(hdr + scan + 2);
    jo_write_headers(real_w, real_h, jo_qtable_luma, jo_qtable_chroma, tables, buf);
    jo_bits_write_stuffed(&bits, buf + hdr);
    buf[hdr + scan] = 0xFF; buf[hdr + scan + 1] = 0xD9;                    /* sEOI */
    jo_bits_free(&bits);
    *out = buf; *out_n = hdr + scan + 2;
    return 0;
}

int jo_encode_rgb(const uint8_t* rgb, uint32_t real_w, uint32_t real_h, uint8_t** out, size_t* out_n) {
    return encode_common(rgb, real_w, real_h, 255, out, out_n);
}

int jo_encode_planes(const double* p0, const double* p1, const double* p2, uint32_t W16, uint32_t H16, uint32_t real_w,
                     uint32_t real_h, int ycbcr, uint8_t** out, size_t* out_n) {
    if (W16 != jo_pad16(real_w) || H16 != jo_pad16(real_h)) return -1;
    const uint32_t mcu_w = W16 / 16, mcu_h = H16 / 16;
    int16_t* coef = (int16_t*)malloc(sizeof(int16_t) * (size_t)mcu_w * mcu_h * 6 * 64);
    jo_forward_from_planes(p0, p1, p2, W16, H16, ycbcr, jo_qtable_luma, jo_qtable_chroma, coef);
    jo_huff_table tables[4];
    jo_bits bits;
    jo_bits_init(&bits);
    jo_entropy_encode(coef, mcu_w, mcu_h, tables, &bits);
    free(coef);
    const size_t hdr = jo_write_headers(real_w, real_h, jo_qtable_luma, jo_qtable_chroma, tables, NULL);
    const size_t scan = jo_bits_stuffed_size(&bits);
    uint8_t* buf = (uint8_t*)malloc(hdr + scan + 2);
    jo_write_headers(real_w, real_h, jo_qtable_luma, jo_qtable_chroma, tables, buf);
    jo_bits_write_stuffed(&bits, buf + hdr);
    buf[hdr + scan] = 0xFF; buf[hdr + scan + 1] = 0xD9;                    /* sEOI */
    jo_bits_free(&bits);
    *out = buf; *out_n = hdr + scan + 2;
    return 0;
}

int jo_encode_ppm(const uint8_t* file, size_t n, uint8_t** out, size_t* out_n) {
    jo_ppm_header h;
    int rc = jo_ppm_parse(file, n, &h);
    if (rc) return rc;
    uint8_t* rgb = (uint8_t*)malloc((size_t)h.width * h.height * 3 + 1);
    rc = jo_ppm_samples(file, n, &h, rgb);
    if (!rc) rc = encode_common(rgb, h.width, h.height, h.maxval, out, out_n);
    free(rgb);
    return rc;
}
