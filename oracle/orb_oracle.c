/*
 * orb_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See orb_oracle.h.
 *
 * Plain C11, no dependencies.  Build with -ffp-contract=off: every float expression below
 * must round once per operation (DESIGN.md parity rule 3, "no FMA").
 *
 * Each function cites the reference lines it restates (paths under /root/reference).
 *
 * PINNED: against the reference's own code compiled in place into oracle/_ref (src/ORBextractor.cc unmodified, the
 * Frame / ORBmatcher hot-path functions and the vendored DBoW2 lifted verbatim; tests/test_oracle_reference.py, test_oracle_bow.py),
 * against OpenCV 4.13 itself through cv2 for the primitives and for cv::ORB (tests/test_oracle_cv2.py, test_oracle_cvorb.py), and by
 * randomised sweeps of both (tools/ref_fuzz.py, ref_fuzz_frame.py, cvorb_fuzz.py; results in profiles/r02_*fuzz*.txt).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; nothing under pointslot_b200/ does.
 */
#include "orb_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PATCH_SIZE 31      /* ORBextractor.cc:72 */
#define HALF_PATCH_SIZE 15 /* ORBextractor.cc:73 */
#define EDGE_THRESHOLD 19  /* ORBextractor.cc:74 */
#define MAX_LEVELS 32

static const int8_t k_pattern[1024] = {
#include "../include/pslot_orb_pattern.inc"
};

/* ------------------------------------------------------------------------------------------ */
/* OpenCV scalar helpers                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* cvRound: round-half-to-even of a double (SSE2 cvtsd2si under the default rounding mode). */
int orc_cv_round(double v) { return (int)lrint(v); }

/* cv::fastAtan2 (OpenCV 4.13 core/mathfuncs_core: atanImpl<float>), degrees in [0,360).
 * float32, one rounding per operation.  Called at ORBextractor.cc:103. */
float orc_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)2.2204460492503131e-16; /* (float)DBL_EPSILON */
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* cv::cvtColor(src, dst, CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) for CV_8U (Tracking.cc:1014-1039):
 * OpenCV 4.13 RGB2Gray<uchar>, 15-bit fixed point, BT.601. */
void orc_cvt_gray(const uint8_t *src, int w, int h, ptrdiff_t spitch, int channels, int rgb_order, uint8_t *dst, ptrdiff_t dpitch)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (ptrdiff_t)y * spitch;
        uint8_t *d = dst + (ptrdiff_t)y * dpitch;
        for (int x = 0; x < w; ++x, s += channels) {
            const unsigned r = rgb_order ? s[0] : s[2], g = s[1], b = rgb_order ? s[2] : s[0];
            d[x] = (uint8_t)((r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* cv::resize(..., INTER_LINEAR) for CV_8UC1 (called at ORBextractor.cc:1120)                 */
/* fixed point: 11-bit coefficients, horizontal pass in int32, vertical pass as in             */
/* OpenCV's VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>.                          */
/* ------------------------------------------------------------------------------------------ */
static void resize_axis_table(int n_dst, int n_src, int *idx0, int *idx1, int16_t *c0, int16_t *c1)
{
    double scale = 1.0 / ((double)n_dst / (double)n_src);
    for (int d = 0; d < n_dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
        idx0[d] = s;
        idx1[d] = s + 1 < n_src ? s + 1 : n_src - 1;
        c0[d] = (int16_t)orc_cv_round((double)((1.f - f) * 2048.f));
        c1[d] = (int16_t)orc_cv_round((double)(f * 2048.f));
    }
}

void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, ptrdiff_t spitch,
                          uint8_t *dst, int dw, int dh, ptrdiff_t dpitch)
{
    int *x0 = malloc(sizeof(int) * (size_t)dw), *x1 = malloc(sizeof(int) * (size_t)dw);
    int *y0 = malloc(sizeof(int) * (size_t)dh), *y1 = malloc(sizeof(int) * (size_t)dh);
    int16_t *cx0 = malloc(2 * (size_t)dw), *cx1 = malloc(2 * (size_t)dw);
    int16_t *cy0 = malloc(2 * (size_t)dh), *cy1 = malloc(2 * (size_t)dh);
    int32_t *h0 = malloc(sizeof(int32_t) * (size_t)dw), *h1 = malloc(sizeof(int32_t) * (size_t)dw);
    resize_axis_table(dw, sw, x0, x1, cx0, cx1);
    resize_axis_table(dh, sh, y0, y1, cy0, cy1);
    for (int dy = 0; dy < dh; ++dy) {
        const uint8_t *r0 = src + (ptrdiff_t)y0[dy] * spitch;
        const uint8_t *r1 = src + (ptrdiff_t)y1[dy] * spitch;
        for (int dx = 0; dx < dw; ++dx) {
            h0[dx] = r0[x0[dx]] * cx0[dx] + r0[x1[dx]] * cx1[dx];
            h1[dx] = r1[x0[dx]] * cx0[dx] + r1[x1[dx]] * cx1[dx];
        }
        int b0 = cy0[dy], b1 = cy1[dy];
        uint8_t *o = dst + (ptrdiff_t)dy * dpitch;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (h0[dx] >> 4)) >> 16) + ((b1 * (h1[dx] >> 4)) >> 16) + 2) >> 2;
            o[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(x0); free(x1); free(y0); free(y1); free(cx0); free(cx1); free(cy0); free(cy1); free(h0); free(h1);
}

/* ------------------------------------------------------------------------------------------ */
/* cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16 (ORBextractor.cc:809,814)  */
/* ------------------------------------------------------------------------------------------ */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* raw corner score: max over the 16 arcs of 9 contiguous ring pixels of min(d) resp. min(-d),
 * minus 1 (OpenCV cornerScore<16>); pixel is a corner at threshold t  <=>  score >= t. */
static inline int fast_raw_score(const uint8_t *p, const ptrdiff_t *off)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best_hi = INT_MIN, best_lo = INT_MIN;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            if (d[k + j] < mn) mn = d[k + j];
            if (d[k + j] > mx) mx = d[k + j];
        }
        if (mn > best_hi) best_hi = mn;
        if (-mx > best_lo) best_lo = -mx;
    }
    return (best_hi > best_lo ? best_hi : best_lo) - 1;
}

/* scratch must hold w*h bytes */
static int fast9_16_impl(const uint8_t *img, int w, int h, ptrdiff_t pitch, int threshold,
                         uint8_t *score, int cap, int *ox, int *oy, int *oresp)
{
    if (w < 7 || h < 7) return 0;
    ptrdiff_t off[16];
    for (int k = 0; k < 16; ++k) off[k] = (ptrdiff_t)k_ring_dy[k] * pitch + k_ring_dx[k];
    memset(score, 0, (size_t)w * (size_t)h);
    const int t = threshold;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *row = img + (ptrdiff_t)y * pitch;
        uint8_t *srow = score + (size_t)y * (size_t)w;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t *p = row + x;
            int v = p[0];
            /* every arc of 9 contains one pixel of each diametral pair: cheap rejection */
            int a0 = v - p[off[0]], a8 = v - p[off[8]];
            int hi = (a0 > t) | (a8 > t), lo = (a0 < -t) | (a8 < -t);
            if (!(hi | lo)) continue;
            int a4 = v - p[off[4]], a12 = v - p[off[12]];
            hi &= (a4 > t) | (a12 > t);
            lo &= (a4 < -t) | (a12 < -t);
            if (!(hi | lo)) continue;
            int s = fast_raw_score(p, off);
            if (s >= t) srow[x] = (uint8_t)s;
        }
    }
    int n = 0;
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *c = score + (size_t)y * (size_t)w;
        const uint8_t *u = c - w, *d = c + w;
        for (int x = 3; x < w - 3; ++x) {
            int s = c[x];
            if (!s) continue;
            if (s > u[x - 1] && s > u[x] && s > u[x + 1] && s > c[x - 1] && s > c[x + 1] &&
                s > d[x - 1] && s > d[x] && s > d[x + 1]) {
                if (n < cap) { ox[n] = x; oy[n] = y; oresp[n] = s; }
                ++n;
            }
        }
    }
    return n;
}

int orc_fast9_16(const uint8_t *img, int w, int h, ptrdiff_t pitch, int threshold,
                 int cap, int *x, int *y, int *resp)
{
    uint8_t *score = malloc((size_t)(w > 0 ? w : 1) * (size_t)(h > 0 ? h : 1));
    int n = fast9_16_impl(img, w, h, pitch, threshold, score, cap, x, y, resp);
    free(score);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on a continuous CV_8U Mat   */
/* (ORBextractor.cc:1085-1086): fixed-point separable kernel [18 34 48 56 48 34 18]/256,       */
/* single rounding at the end.                                                                 */
/* ------------------------------------------------------------------------------------------ */
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

void orc_gaussian7x7_s2(const uint8_t *src, int w, int h, ptrdiff_t spitch, uint8_t *dst, ptrdiff_t dpitch)
{
    static const int kw[7] = {18, 34, 48, 56, 48, 34, 18};
    int32_t *hbuf = malloc(sizeof(int32_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *r = src + (ptrdiff_t)y * spitch;
        int32_t *o = hbuf + (size_t)y * (size_t)w;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x < w - 3) {
                for (int k = 0; k < 7; ++k) acc += kw[k] * r[x + k - 3];
            } else {
                for (int k = 0; k < 7; ++k) acc += kw[k] * r[reflect101(x + k - 3, w)];
            }
            o[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const int32_t *rows[7];
        for (int k = 0; k < 7; ++k) rows[k] = hbuf + (size_t)reflect101(y + k - 3, h) * (size_t)w;
        uint8_t *o = dst + (ptrdiff_t)y * dpitch;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int k = 0; k < 7; ++k) acc += kw[k] * rows[k][x];
            o[x] = (uint8_t)((acc + 32768) >> 16);
        }
    }
    free(hbuf);
}

/* ------------------------------------------------------------------------------------------ */
/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:2704-2720)                                    */
/* ------------------------------------------------------------------------------------------ */
int orc_hamming256(const void *a, const void *b)
{
    uint32_t pa[8], pb[8];
    memcpy(pa, a, 32);
    memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* ------------------------------------------------------------------------------------------ */
/* DistributeOctTree (ORBextractor.cc:539-763) + ExtractorNode::DivideNode (:481-537)          */
/* ------------------------------------------------------------------------------------------ */
typedef struct qnode {
    int ulx, uly, urx, ury, blx, bly, brx, bry; /* UL, UR, BL, BR corners */
    int *keys;                                  /* indices into the candidate arrays, insertion order */
    int nkeys;
    int nomore;
    long seq; /* creation sequence number: the canonical stand-in for the heap address (rule C.1) */
    struct qnode *prev, *next;
} qnode;

typedef struct { qnode *head, *tail; int size; long next_seq; } qlist;

static qnode *qnode_new(int reserve)
{
    qnode *n = calloc(1, sizeof(qnode));
    n->keys = malloc(sizeof(int) * (size_t)(reserve > 0 ? reserve : 1));
    return n;
}
static void qnode_free(qnode *n) { free(n->keys); free(n); }
static void qlist_push_front(qlist *l, qnode *n)
{
    n->seq = l->next_seq++;
    n->prev = NULL; n->next = l->head;
    if (l->head) l->head->prev = n; else l->tail = n;
    l->head = n; l->size++;
}
static void qlist_push_back(qlist *l, qnode *n)
{
    n->seq = l->next_seq++;
    n->next = NULL; n->prev = l->tail;
    if (l->tail) l->tail->next = n; else l->head = n;
    l->tail = n; l->size++;
}
static qnode *qlist_erase(qlist *l, qnode *n) /* returns the following node */
{
    qnode *nx = n->next;
    if (n->prev) n->prev->next = n->next; else l->head = n->next;
    if (n->next) n->next->prev = n->prev; else l->tail = n->prev;
    l->size--;
    qnode_free(n);
    return nx;
}

/* DivideNode (:481-537).  x,y are the candidate coordinate arrays. */
static void divide_node(const qnode *p, const int *kx, const int *ky, qnode *c[4])
{
    const int halfX = (int)ceilf((float)(p->urx - p->ulx) / 2);
    const int halfY = (int)ceilf((float)(p->bry - p->uly) / 2);
    for (int i = 0; i < 4; ++i) c[i] = qnode_new(p->nkeys);
    qnode *n1 = c[0], *n2 = c[1], *n3 = c[2], *n4 = c[3];
    n1->ulx = p->ulx; n1->uly = p->uly;
    n1->urx = p->ulx + halfX; n1->ury = p->uly;
    n1->blx = p->ulx; n1->bly = p->uly + halfY;
    n1->brx = p->ulx + halfX; n1->bry = p->uly + halfY;

    n2->ulx = n1->urx; n2->uly = n1->ury;
    n2->urx = p->urx; n2->ury = p->ury;
    n2->blx = n1->brx; n2->bly = n1->bry;
    n2->brx = p->urx; n2->bry = p->uly + halfY;

    n3->ulx = n1->blx; n3->uly = n1->bly;
    n3->urx = n1->brx; n3->ury = n1->bry;
    n3->blx = p->blx; n3->bly = p->bly;
    n3->brx = n1->brx; n3->bry = p->bly;

    n4->ulx = n3->urx; n4->uly = n3->ury;
    n4->urx = n2->brx; n4->ury = n2->bry;
    n4->blx = n3->brx; n4->bly = n3->bry;
    n4->brx = p->brx; n4->bry = p->bry;

    for (int i = 0; i < p->nkeys; ++i) {
        int k = p->keys[i];
        qnode *dst;
        if (kx[k] < n1->urx) dst = (ky[k] < n1->bry) ? n1 : n3;
        else dst = (ky[k] < n1->bry) ? n2 : n4;
        dst->keys[dst->nkeys++] = k;
    }
    for (int i = 0; i < 4; ++i)
        if (c[i]->nkeys == 1) c[i]->nomore = 1;
}

typedef struct { int size; long seq; qnode *node; } qsized;
static int qsized_cmp(const void *a, const void *b)
{
    const qsized *pa = a, *pb = b;
    if (pa->size != pb->size) return pa->size < pb->size ? -1 : 1;
    if (pa->seq != pb->seq) return pa->seq < pb->seq ? -1 : 1; /* rule C.1: later-created sorts last */
    return 0;
}

/* push the non-empty children in order n1..n4, recording expandable ones; frees empty ones */
static void push_children(qlist *l, qnode *c[4], qsized **v, int *nv, int *capv, int *n_to_expand)
{
    for (int i = 0; i < 4; ++i) {
        if (c[i]->nkeys > 0) {
            qlist_push_front(l, c[i]);
            if (c[i]->nkeys > 1) {
                if (n_to_expand) (*n_to_expand)++;
                if (*nv == *capv) { *capv = *capv * 2 + 16; *v = realloc(*v, sizeof(qsized) * (size_t)*capv); }
                (*v)[*nv].size = c[i]->nkeys; (*v)[*nv].seq = c[i]->seq; (*v)[*nv].node = c[i];
                (*nv)++;
            }
        } else {
            qnode_free(c[i]);
        }
    }
}

int orc_distribute_octtree(const int *x, const int *y, const int *resp, int n,
                           int minX, int maxX, int minY, int maxY, int N,
                           int cap, int *ox, int *oy, int *oresp, int *tie_sensitive)
{
    if (tie_sensitive) *tie_sensitive = 0;
    /* :543-545 */
    const int nIni = (int)roundf((float)(maxX - minX) / (float)(maxY - minY));
    const float hX = (float)(maxX - minX) / (float)nIni;
    if (nIni <= 0) return 0; /* reference divides by zero here; such shapes are rejected upstream */

    qlist L = {0};
    qnode **roots = malloc(sizeof(qnode *) * (size_t)nIni);
    for (int i = 0; i < nIni; ++i) { /* :550-563 */
        qnode *ni = qnode_new(n);
        ni->ulx = (int)(hX * (float)i); ni->uly = 0;
        ni->urx = (int)(hX * (float)(i + 1)); ni->ury = 0;
        ni->blx = ni->ulx; ni->bly = maxY - minY;
        ni->brx = ni->urx; ni->bry = maxY - minY;
        qlist_push_back(&L, ni);
        roots[i] = ni;
    }
    /* push_back'ed roots must order front->back as 0..nIni-1 while every later push_front lands
       before them; seq is only used for sorting ties among *children*, so the root values are moot */
    for (int i = 0; i < n; ++i) { /* :566-570 */
        int r = (int)((float)x[i] / hX);
        if (r >= nIni) r = nIni - 1; /* cannot happen for in-window keys; guards the oracle only */
        roots[r]->keys[roots[r]->nkeys++] = i;
    }
    free(roots);
    for (qnode *it = L.head; it;) { /* :572-585 */
        if (it->nkeys == 1) { it->nomore = 1; it = it->next; }
        else if (it->nkeys == 0) it = qlist_erase(&L, it);
        else it = it->next;
    }

    int finish = 0;
    qsized *V = NULL; int nV = 0, capV = 0;
    while (!finish) { /* :594 */
        int prevSize = L.size;
        int nToExpand = 0;
        nV = 0;
        /* coarse pass (:600-665): visit only the nodes present at the start of the pass */
        qnode *it = L.head;
        while (it) {
            if (it->nomore) { it = it->next; continue; }
            qnode *c[4];
            divide_node(it, x, y, c);
            push_children(&L, c, &V, &nV, &capV, &nToExpand);
            it = qlist_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) { /* :669 */
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) { /* :673 */
            while (!finish) {
                prevSize = L.size;
                int nP = nV;
                qsized *P = malloc(sizeof(qsized) * (size_t)(nP > 0 ? nP : 1));
                memcpy(P, V, sizeof(qsized) * (size_t)nP);
                nV = 0;
                qsort(P, (size_t)nP, sizeof(qsized), qsized_cmp); /* :684 */
                for (int j = nP - 1; j >= 0; --j) {
                    qnode *c[4];
                    divide_node(P[j].node, x, y, c);
                    push_children(&L, c, &V, &nV, &capV, NULL);
                    qlist_erase(&L, P[j].node);
                    /* Rule C.1 bookkeeping.  bit 1: two equal-size nodes are split in this run - under another allocator
                       they are split in another order, which permutes the list (keypoint order, descriptor rows, stereo iR
                       ties) without changing the set.  bit 0: the cut of :731 comes at a node that belongs to a run of
                       equal-size nodes - the retained SET depends on the tie rule: either equal-size nodes stay unsplit
                       behind the cut, or (cut at the end of the run) another order inside the run reaches N earlier,
                       because a split adds 0 to 3 nodes.  A cut at a node of unique size is safe: every larger run was
                       split completely and would be in any order.  Checked against the stock-allocator build of the
                       reference: a level with flags == 0 is byte-identical, one without bit 0 holds the same set. */
                    const int next_equal = j + 1 < nP && P[j + 1].size == P[j].size;
                    if (tie_sensitive && next_equal && L.size < N) *tie_sensitive |= 2;
                    if (L.size >= N) { /* :731 */
                        if (tie_sensitive && (next_equal || (j > 0 && P[j - 1].size == P[j].size))) *tie_sensitive |= 1;
                        break;
                    }
                }
                free(P);
                if (L.size >= N || L.size == prevSize) finish = 1; /* :735 */
            }
        }
    }
    free(V);

    /* :741-760 keep the first maximum-response key of each node, list order */
    int no = 0;
    for (qnode *it = L.head; it;) {
        int best = it->keys[0];
        for (int k = 1; k < it->nkeys; ++k)
            if ((float)resp[it->keys[k]] > (float)resp[best]) best = it->keys[k];
        if (no < cap) { ox[no] = x[best]; oy[no] = y[best]; oresp[no] = resp[best]; }
        ++no;
        qnode *nx = it->next;
        qnode_free(it);
        it = nx;
    }
    return no;
}

/* ------------------------------------------------------------------------------------------ */
/* ORBextractor                                                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int w, h;
    uint8_t *img;     /* borderless level image, pitch = w (the 19-px frame of :1113-1128 is never read) */
    uint8_t *blurred; /* GaussianBlur of img, NULL if the level has no keypoints */
    int ncand, capcand;
    int *cx, *cy, *cresp;
    int nkp, capkp;
    int *kx, *ky, *kresp;
    float *kangle;
    int tie_sensitive;
} orc_level;

struct orc_extractor {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;
    float sf[MAX_LEVELS], inv_sf[MAX_LEVELS], sigma2[MAX_LEVELS], inv_sigma2[MAX_LEVELS];
    int per_level[MAX_LEVELS];
    int umax[HALF_PATCH_SIZE + 1];
    orc_level lv[MAX_LEVELS];
};

orc_extractor *orc_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th)
{
    if (nlevels < 1 || nlevels > MAX_LEVELS || nfeatures < 1) return NULL;
    orc_extractor *e = calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = ini_th; e->min_th = min_th;
    e->scale_factor = (double)scale_factor; /* ORBextractor.h:98 double member, float argument */
    /* :415-431 */
    e->sf[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        e->sf[i] = (float)((double)e->sf[i - 1] * e->scale_factor);
        e->sigma2[i] = e->sf[i] * e->sf[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        e->inv_sf[i] = 1.0f / e->sf[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    /* :435-446 */
    float factor = (float)(1.0 / e->scale_factor);
    float nDesired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e->per_level[l] = orc_cv_round((double)nDesired);
        sum += e->per_level[l];
        nDesired *= factor;
    }
    e->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* :454-469 */
    int v, v0;
    int vmax = (int)floor((double)(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1));
    int vmin = (int)ceil((double)(HALF_PATCH_SIZE * sqrtf(2.f) / 2));
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = orc_cv_round(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

static void level_release(orc_level *L)
{
    free(L->img); free(L->blurred);
    free(L->cx); free(L->cy); free(L->cresp);
    free(L->kx); free(L->ky); free(L->kresp); free(L->kangle);
    memset(L, 0, sizeof(*L));
}

void orc_destroy(orc_extractor *e)
{
    if (!e) return;
    for (int l = 0; l < e->nlevels; ++l) level_release(&e->lv[l]);
    free(e);
}

const float *orc_scale_factors(const orc_extractor *e) { return e->sf; }
const float *orc_inv_scale_factors(const orc_extractor *e) { return e->inv_sf; }
const float *orc_level_sigma2(const orc_extractor *e) { return e->sigma2; }
const float *orc_inv_level_sigma2(const orc_extractor *e) { return e->inv_sigma2; }
const int *orc_features_per_level(const orc_extractor *e) { return e->per_level; }
const int *orc_umax(const orc_extractor *e) { return e->umax; }

/* ComputePyramid (:1107-1132) */
static void compute_pyramid(orc_extractor *e, const uint8_t *img, int w, int h, ptrdiff_t pitch)
{
    for (int l = 0; l < e->nlevels; ++l) {
        orc_level *L = &e->lv[l];
        float scale = e->inv_sf[l];
        int lw = orc_cv_round((double)((float)w * scale));
        int lh = orc_cv_round((double)((float)h * scale));
        free(L->img); free(L->blurred); L->blurred = NULL;
        L->w = lw; L->h = lh;
        L->img = malloc((size_t)(lw > 0 ? lw : 1) * (size_t)(lh > 0 ? lh : 1));
        if (l == 0) {
            for (int y = 0; y < lh; ++y) memcpy(L->img + (size_t)y * (size_t)lw, img + (ptrdiff_t)y * pitch, (size_t)lw);
        } else {
            orc_level *P = &e->lv[l - 1];
            orc_resize_linear_u8(P->img, P->w, P->h, P->w, L->img, lw, lh, lw);
        }
    }
}

static void cand_push(orc_level *L, int x, int y, int r)
{
    if (L->ncand == L->capcand) {
        L->capcand = L->capcand * 2 + 1024;
        L->cx = realloc(L->cx, sizeof(int) * (size_t)L->capcand);
        L->cy = realloc(L->cy, sizeof(int) * (size_t)L->capcand);
        L->cresp = realloc(L->cresp, sizeof(int) * (size_t)L->capcand);
    }
    L->cx[L->ncand] = x; L->cy[L->ncand] = y; L->cresp[L->ncand] = r; L->ncand++;
}

/* IC_Angle (:77-104) */
static float ic_angle(const uint8_t *img, int pitch, int px, int py, const int *umax)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (ptrdiff_t)py * pitch + px;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * pitch], val_minus = center[u - v * pitch];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* ComputeKeyPointsOctTree (:765-853) */
static int compute_keypoints_octtree(orc_extractor *e)
{
    const float W = 30;
    uint8_t *score = NULL; size_t score_cap = 0;
    int *fx = NULL, *fy = NULL, *fr = NULL; int fcap = 0;
    for (int level = 0; level < e->nlevels; ++level) {
        orc_level *L = &e->lv[level];
        L->ncand = 0; L->nkp = 0; L->tie_sensitive = 0;
        const int minBorderX = EDGE_THRESHOLD - 3;
        const int minBorderY = minBorderX;
        const int maxBorderX = L->w - EDGE_THRESHOLD + 3;
        const int maxBorderY = L->h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBorderX - minBorderX);
        const float height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W);
        const int nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) { free(score); free(fx); free(fy); free(fr); return -2; } /* reference: /0 */
        const int wCell = (int)ceilf(width / (float)nCols);
        const int hCell = (int)ceilf(height / (float)nRows);

        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + (float)hCell + 6;
            if (iniY >= (float)(maxBorderY - 3)) continue;
            if (maxY > (float)maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + (float)wCell + 6;
                if (iniX >= (float)(maxBorderX - 3)) continue; /* :803 (upstream has -6) */
                if (maxX > (float)maxBorderX) maxX = (float)maxBorderX;
                /* rowRange(iniY,maxY).colRange(iniX,maxX): float -> int truncation */
                int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
                int cw = x1 - x0, ch = y1 - y0;
                if ((size_t)cw * (size_t)ch > score_cap) { score_cap = (size_t)cw * (size_t)ch; score = realloc(score, score_cap); }
                int maxk = ((cw + 1) / 2) * ((ch + 1) / 2) + 1;
                if (maxk > fcap) { fcap = maxk; fx = realloc(fx, sizeof(int) * (size_t)fcap); fy = realloc(fy, sizeof(int) * (size_t)fcap); fr = realloc(fr, sizeof(int) * (size_t)fcap); }
                const uint8_t *sub = L->img + (ptrdiff_t)y0 * L->w + x0;
                int nk = fast9_16_impl(sub, cw, ch, L->w, e->ini_th, score, fcap, fx, fy, fr);
                if (nk == 0) nk = fast9_16_impl(sub, cw, ch, L->w, e->min_th, score, fcap, fx, fy, fr);
                for (int k = 0; k < nk; ++k) cand_push(L, fx[k] + j * wCell, fy[k] + i * hCell, fr[k]);
            }
        }

        int capk = e->per_level[level] + 8;
        if (capk < L->ncand + 1) capk = L->ncand + 1; /* list can never exceed the candidate count */
        if (capk > L->capkp) {
            L->capkp = capk;
            L->kx = realloc(L->kx, sizeof(int) * (size_t)capk);
            L->ky = realloc(L->ky, sizeof(int) * (size_t)capk);
            L->kresp = realloc(L->kresp, sizeof(int) * (size_t)capk);
            L->kangle = realloc(L->kangle, sizeof(float) * (size_t)capk);
        }
        /* :543 nIni = round(width / height): 0 for portrait levels narrower than half their height, where the
           reference divides by zero and indexes an empty vector -> rejected like the too-small case above */
        if ((int)roundf((float)(maxBorderX - minBorderX) / (float)(maxBorderY - minBorderY)) < 1) {
            free(score); free(fx); free(fy); free(fr);
            return -2;
        }
        L->nkp = orc_distribute_octtree(L->cx, L->cy, L->cresp, L->ncand, minBorderX, maxBorderX,
                                        minBorderY, maxBorderY, e->per_level[level],
                                        L->capkp, L->kx, L->ky, L->kresp, &L->tie_sensitive);
        for (int i = 0; i < L->nkp; ++i) { /* :837-847 */
            L->kx[i] += minBorderX;
            L->ky[i] += minBorderY;
        }
    }
    free(score); free(fx); free(fy); free(fr);
    for (int level = 0; level < e->nlevels; ++level) { /* :850-852 */
        orc_level *L = &e->lv[level];
        for (int i = 0; i < L->nkp; ++i) L->kangle[i] = ic_angle(L->img, L->w, L->kx[i], L->ky[i], e->umax);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* cosf / sinf exactly as the reference binary gets them.  computeOrbDescriptor writes          */
/* `(float)cos(angle)` with a float argument under `using namespace std` (ORBextractor.cc:      */
/* 112-113), which resolves to std::cos(float) = glibc cosf.  glibc >= 2.28 implements          */
/* sinf/cosf/sincosf with the algorithm of ARM's optimized-routines (sysdeps/ieee754/flt-32/    */
/* s_sincosf.h): reduce by pi/2 with a 2^24-prescaled multiply, then an even or odd polynomial  */
/* in double, one final rounding to float.  Restated here (constants = glibc's                  */
/* __sincosf_table) so the oracle does not depend on the host libm; tests/test_oracle_cv2.py    */
/* checks it against this image's libm (glibc 2.39), and an exhaustive sweep of all 1.09e9      */
/* floats in [0, 2 pi] found no differing result for cosf nor sinf, with or without FMA          */
/* contraction of the double arithmetic (profiles/r02_sincosf_sweep.txt).  Only the ranges the  */
/* path can reach are covered: |x| < 120.                                                       */
/* ------------------------------------------------------------------------------------------ */
static const double SC_HPI_INV = 0x1.45F306DC9C883p+23, SC_HPI = 0x1.921FB54442D18p0;
static const double SC_C[5] = {0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16};
static const double SC_S[3] = {-0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13};
static inline uint32_t sc_abstop12(float x) { uint32_t u; memcpy(&u, &x, 4); return (u >> 20) & 0x7ff; }
static inline float sc_poly(double x, double x2, int negcos, int n)
{
    if ((n & 1) == 0) { /* sine polynomial */
        const double x3 = x * x2, s1 = SC_S[1] + x2 * SC_S[2], x7 = x3 * x2, s = x + x3 * SC_S[0];
        return (float)(s + x7 * s1);
    }
    const double g = negcos ? -1.0 : 1.0; /* glibc's second table row: the cosine coefficients negated */
    const double x4 = x2 * x2, c2 = g * SC_C[3] + x2 * (g * SC_C[4]), c1 = g * SC_C[0] + x2 * (g * SC_C[1]);
    const double x6 = x4 * x2, c = c1 + x4 * (g * SC_C[2]);
    return (float)(c + x6 * c2);
}
static inline float sc_eval(float y, int cosine)
{
    static const double sign[4] = {1.0, -1.0, -1.0, 1.0};
    double x = y;
    if (sc_abstop12(y) < sc_abstop12(0x1.921FB6p-1f)) { /* |y| < pi/4 */
        if (sc_abstop12(y) < sc_abstop12(0x1p-12f)) return cosine ? 1.0f : y;
        return sc_poly(x, x * x, 0, cosine);
    }
    /* reduce_fast: valid for |y| < 120 (angles here are below 2 pi) */
    const double r = x * SC_HPI_INV;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * SC_HPI;
    return sc_poly(x * sign[n & 3], x * x, (n & 2) != 0, cosine ? n ^ 1 : n);
}
float orc_glibc_cosf(float x) { return sc_eval(x, 1); }
float orc_glibc_sinf(float x) { return sc_eval(x, 0); }
void orc_glibc_sincosf_n(const float *x, int n, float *c, float *s)
{
    for (int i = 0; i < n; ++i) { c[i] = sc_eval(x[i], 1); s[i] = sc_eval(x[i], 0); }
}

/* steering coefficients (:112-113): cosf / sinf of the float angle (DESIGN.md parity rule 3) */
static inline void steer_coeffs(float kp_angle, float *a, float *b)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = kp_angle * factorPI;
    *a = orc_glibc_cosf(angle);
    *b = orc_glibc_sinf(angle);
}

void orc_steer_coeffs(const float *angle_deg, int n, float *a, float *b)
{
    for (int i = 0; i < n; ++i) steer_coeffs(angle_deg[i], &a[i], &b[i]);
}

/* computeOrbDescriptor (:107-147) */
static void compute_orb_descriptor(const uint8_t *img, int step, int px, int py, float kp_angle, uint8_t *desc)
{
    float a, b;
    steer_coeffs(kp_angle, &a, &b);
    const uint8_t *center = img + (ptrdiff_t)py * step + px;
    const int8_t *pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int t = 0; t < 8; ++t) {
            const int8_t *q = pat + 4 * t;
            float x0 = (float)q[0], y0 = (float)q[1], x1 = (float)q[2], y1 = (float)q[3];
            int t0 = center[orc_cv_round((double)(x0 * b + y0 * a)) * step + orc_cv_round((double)(x0 * a - y0 * b))];
            int t1 = center[orc_cv_round((double)(x1 * b + y1 * a)) * step + orc_cv_round((double)(x1 * a - y1 * b))];
            val |= (t0 < t1) << t;
        }
        desc[i] = (uint8_t)val;
    }
}

int orc_extract(orc_extractor *e, const uint8_t *img, int w, int h, ptrdiff_t pitch,
                orc_keypoint *kps, uint8_t *desc32, int cap, int *n)
{
    *n = 0;
    if (!img || w <= 0 || h <= 0) return 0; /* :1046-1047 */
    compute_pyramid(e, img, w, h, pitch);
    int rc = compute_keypoints_octtree(e);
    if (rc) return rc;
    int total = 0;
    for (int l = 0; l < e->nlevels; ++l) total += e->lv[l].nkp;
    *n = total;
    if (total > cap) return -1;
    int offset = 0;
    for (int l = 0; l < e->nlevels; ++l) { /* :1075-1104 */
        orc_level *L = &e->lv[l];
        if (L->nkp == 0) continue;
        L->blurred = malloc((size_t)L->w * (size_t)L->h);
        orc_gaussian7x7_s2(L->img, L->w, L->h, L->w, L->blurred, L->w);
        const int scaledPatchSize = (int)((float)PATCH_SIZE * e->sf[l]); /* :834 */
        for (int i = 0; i < L->nkp; ++i) {
            orc_keypoint *k = &kps[offset + i];
            compute_orb_descriptor(L->blurred, L->w, L->kx[i], L->ky[i], L->kangle[i], desc32 + 32 * (size_t)(offset + i));
            k->x = (float)L->kx[i]; k->y = (float)L->ky[i];
            if (l != 0) { k->x *= e->sf[l]; k->y *= e->sf[l]; } /* :1095-1101 */
            k->size = (float)scaledPatchSize;
            k->angle = L->kangle[i];
            k->response = (float)L->kresp[i];
            k->octave = l;
            k->class_id = -1;
        }
        offset += L->nkp;
    }
    return 0;
}

int orc_level_size(const orc_extractor *e, int level, int *w, int *h)
{
    if (level < 0 || level >= e->nlevels) return -1;
    *w = e->lv[level].w; *h = e->lv[level].h;
    return 0;
}
const uint8_t *orc_level_ptr(const orc_extractor *e, int level) { return e->lv[level].img; }
const uint8_t *orc_blurred_ptr(const orc_extractor *e, int level) { return e->lv[level].blurred; }
int orc_level_candidates(const orc_extractor *e, int level, int cap, int *x, int *y, int *resp)
{
    const orc_level *L = &e->lv[level];
    for (int i = 0; i < L->ncand && i < cap; ++i) { x[i] = L->cx[i]; y[i] = L->cy[i]; resp[i] = L->cresp[i]; }
    return L->ncand;
}
int orc_level_keypoints(const orc_extractor *e, int level, int cap, int *x, int *y, int *resp, float *angle)
{
    const orc_level *L = &e->lv[level];
    for (int i = 0; i < L->nkp && i < cap; ++i) { x[i] = L->kx[i]; y[i] = L->ky[i]; resp[i] = L->kresp[i]; angle[i] = L->kangle[i]; }
    return L->nkp;
}
int orc_level_tie_sensitive(const orc_extractor *e, int level) { return e->lv[level].tie_sensitive; }

/* ------------------------------------------------------------------------------------------ */
/* Frame::ComputeStereoMatches (Frame.cc:2142-2316)                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int dist, idx; } dist_idx;
static int dist_idx_cmp(const void *a, const void *b)
{
    const dist_idx *pa = a, *pb = b;
    if (pa->dist != pb->dist) return pa->dist < pb->dist ? -1 : 1;
    if (pa->idx != pb->idx) return pa->idx < pb->idx ? -1 : 1;
    return 0;
}

int orc_stereo_match(const orc_extractor *eL, const orc_extractor *eR,
                     const orc_keypoint *kL, int nL, const uint8_t *dL,
                     const orc_keypoint *kR, int nR, const uint8_t *dR,
                     float bf, float maxD, float *uright, float *depth,
                     int *best_dist_out, int *best_idx_out, int *best_sad_out)
{
    const int TH_HIGH = 100, TH_LOW = 50; /* ORBmatcher.cc:58-62 */
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    for (int i = 0; i < nL; ++i) {
        uright[i] = -1.0f; depth[i] = -1.0f;
        if (best_dist_out) best_dist_out[i] = TH_HIGH;
        if (best_idx_out) best_idx_out[i] = -1;
        if (best_sad_out) best_sad_out[i] = -1;
    }
    if (nL == 0) return 0;
    const int nRows = eL->lv[0].h;
    /* row table (:2152-2169) */
    int *rcount = calloc((size_t)nRows + 1, sizeof(int));
    for (int pass = 0; pass < 1; ++pass)
        for (int iR = 0; iR < nR; ++iR) {
            const float kpY = kR[iR].y;
            const float r = 2.0f * eL->sf[kR[iR].octave];
            const int maxr = (int)ceilf(kpY + r);
            const int minr = (int)floorf(kpY - r);
            for (int yi = minr; yi <= maxr; ++yi)
                if (yi >= 0 && yi < nRows) rcount[yi]++;
        }
    int *rstart = malloc(sizeof(int) * ((size_t)nRows + 1));
    int tot = 0;
    for (int i = 0; i < nRows; ++i) { rstart[i] = tot; tot += rcount[i]; rcount[i] = 0; }
    rstart[nRows] = tot;
    int *rlist = malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
    for (int iR = 0; iR < nR; ++iR) {
        const float kpY = kR[iR].y;
        const float r = 2.0f * eL->sf[kR[iR].octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) rlist[rstart[yi] + rcount[yi]++] = iR;
    }

    const float minD = 0;
    dist_idx *vDistIdx = malloc(sizeof(dist_idx) * (size_t)nL);
    int nDist = 0;

    for (int iL = 0; iL < nL; ++iL) { /* :2180 */
        const orc_keypoint *kpL = &kL[iL];
        const int levelL = kpL->octave;
        const float vL = kpL->y, uL = kpL->x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const int nc = rcount[row];
        if (nc == 0) continue;
        const int *cands = rlist + rstart[row];
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        const uint8_t *dl = dL + 32 * (size_t)iL;
        for (int iC = 0; iC < nc; ++iC) {
            const int iR = cands[iC];
            const orc_keypoint *kpR = &kR[iR];
            if (kpR->octave < levelL - 1 || kpR->octave > levelL + 1) continue;
            const float uR = kpR->x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_hamming256(dl, dR + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (best_dist_out) best_dist_out[iL] = bestDist;
        if (best_idx_out && bestDist < TH_HIGH) best_idx_out[iL] = bestIdxR;

        if (bestDist < thOrbDist) { /* :2228 */
            const float uR0 = kR[bestIdxR].x;
            const float scaleFactor = eL->inv_sf[kpL->octave];
            const float scaleduL = roundf(kpL->x * scaleFactor);
            const float scaledvL = roundf(kpL->y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const orc_level *PL = &eL->lv[kpL->octave], *PR = &eR->lv[kpL->octave];
            const int ly0 = (int)(scaledvL - w), lx0 = (int)(scaleduL - w);
            int bestSAD = INT_MAX;
            int bestincR = 0;
            const int Lh = 5;
            float vDists[11];
            const float iniu = scaleduR0 + Lh - w;
            const float endu = scaleduR0 + Lh + w + 1;
            if (iniu < 0 || endu >= (float)PR->w) continue;
            const int cl = PL->img[(size_t)(ly0 + w) * (size_t)PL->w + (size_t)(lx0 + w)];
            for (int incR = -Lh; incR <= Lh; ++incR) {
                const int rx0 = (int)(scaleduR0 + (float)incR - w);
                const int cr = PR->img[(size_t)(ly0 + w) * (size_t)PR->w + (size_t)(rx0 + w)];
                /* cv::norm(IL,IR,NORM_L1) of the centre-subtracted float patches: integer-exact */
                int acc = 0;
                for (int yy = 0; yy < 2 * w + 1; ++yy) {
                    const uint8_t *pl = PL->img + (size_t)(ly0 + yy) * (size_t)PL->w + lx0;
                    const uint8_t *pr = PR->img + (size_t)(ly0 + yy) * (size_t)PR->w + rx0;
                    for (int xx = 0; xx < 2 * w + 1; ++xx) acc += abs((pl[xx] - cl) - (pr[xx] - cr));
                }
                float dist = (float)acc;
                if (dist < (float)bestSAD) { bestSAD = (int)dist; bestincR = incR; }
                vDists[Lh + incR] = dist;
            }
            if (bestincR == -Lh || bestincR == Lh) continue;
            const float dist1 = vDists[Lh + bestincR - 1];
            const float dist2 = vDists[Lh + bestincR];
            const float dist3 = vDists[Lh + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = eL->sf[kpL->octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = (float)0.01;
                    bestuR = (float)((double)uL - 0.01);
                }
                depth[iL] = bf / disparity;
                uright[iL] = bestuR;
                if (best_sad_out) best_sad_out[iL] = bestSAD;
                vDistIdx[nDist].dist = bestSAD; vDistIdx[nDist].idx = iL; nDist++;
            }
        }
    }
    if (nDist > 0) { /* :2302-2315 (empty -> no-op, rule C.4) */
        qsort(vDistIdx, (size_t)nDist, sizeof(dist_idx), dist_idx_cmp);
        const float median = (float)vDistIdx[nDist / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nDist - 1; i >= 0; --i) {
            if ((float)vDistIdx[i].dist < thDist) break;
            uright[vDistIdx[i].idx] = -1;
            depth[vDistIdx[i].idx] = -1;
        }
    }
    free(vDistIdx); free(rlist); free(rstart); free(rcount);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Feature grid: Frame::PosInGrid (Frame.cc:2027-2038), Frame::AssignFeaturesToGrid
 * (Frame.cc:1636-1655) and its consumer Frame::GetFeaturesInArea (Frame.cc:1808-1861).
 * mGrid[i][j] is returned in CSR form: cell = i * rows + j, cell_start[cols * rows + 1], indices in push_back
 * order (ascending keypoint index inside a cell).
 * ------------------------------------------------------------------------------------------------------------ */
static int pos_in_grid(const orc_keypoint *kp, float min_x, float min_y, float inv_w, float inv_h, int cols, int rows)
{
    /* posX = round((kp.pt.x-mnMinX)*mfGridElementWidthInv): float product, C round() on its double promotion */
    const double rx = round((double)((kp->x - min_x) * inv_w));
    const double ry = round((double)((kp->y - min_y) * inv_h));
    /* the reference converts to int before the range test; the float test is the same for every finite value
     * that fits an int, and rejects what would be undefined behaviour there */
    if (!(rx >= 0 && rx < cols && ry >= 0 && ry < rows)) return -1;
    return (int)rx * rows + (int)ry;
}

int orc_assign_grid(const orc_keypoint *kps, int n, float min_x, float min_y, float inv_w, float inv_h, int cols,
                    int rows, int *cell_start, int *indices)
{
    const int C = cols * rows;
    if (n < 0 || cols < 1 || rows < 1) return -1;
    memset(cell_start, 0, sizeof(int) * (size_t)(C + 1));
    for (int i = 0; i < n; ++i) {
        const int c = pos_in_grid(&kps[i], min_x, min_y, inv_w, inv_h, cols, rows);
        if (c >= 0) cell_start[c + 1]++;
    }
    for (int c = 0; c < C; ++c) cell_start[c + 1] += cell_start[c];
    int *fill = (int *)malloc(sizeof(int) * (size_t)C);
    memcpy(fill, cell_start, sizeof(int) * (size_t)C);
    for (int i = 0; i < n; ++i) { /* ascending i == push_back order (:1647-1654) */
        const int c = pos_in_grid(&kps[i], min_x, min_y, inv_w, inv_h, cols, rows);
        if (c >= 0) indices[fill[c]++] = i;
    }
    free(fill);
    return cell_start[C];
}

/* Frame::AssignFeatures (Frame.cc:811-905): the two classification loops (ORB features :813-868, dense object features
 * :870-903) differ only in what they do with the background class, so one routine returns the class of every feature
 * and the push_back order per class. */
int orc_assign_features(const orc_keypoint *kps, int n, const uint8_t *mask, int w, int h, int pitch, const int *object_ids,
                        int n_objects, const float *depth, float depth_limit, int *cls, int *order, int *class_start,
                        int *effective)
{
    if (n < 0 || n_objects < 0 || w <= 0 || h <= 0 || pitch < w) return -1;
    memset(class_start, 0, sizeof(int) * (size_t)(n_objects + 2));
    if (effective) memset(effective, 0, sizeof(int) * (size_t)n_objects);
    for (int i = 0; i < n; ++i) {
        const int x = (int)kps[i].x, y = (int)kps[i].y; /* int x = mvKeys[i].pt.x (:816-817) */
        int c = -1;
        if (x >= 0 && y >= 0 && x < w && y < h) {
            const int label = mask[(size_t)y * pitch + x];
            if (label == 0)
                c = 0; /* :820-829 */
            else if (label != 255) { /* :835 */
                for (int j = 0; j < n_objects; ++j) { /* :837-866: first match wins (break) */
                    if (object_ids[j] < 0) continue;  /* if (!vDetectionObjects[j]) continue; negative ids match no label either */
                    const int id = object_ids[j] > 255 ? object_ids[j] - 255 : object_ids[j];
                    if (id == label - 1) { c = 1 + j; break; }
                }
            }
        }
        cls[i] = c;
        if (c >= 0) class_start[c + 1]++;
        if (c >= 1 && depth && effective && depth[i] > 0 && depth[i] < depth_limit) effective[c - 1]++; /* :893-896 */
    }
    for (int c = 0; c <= n_objects; ++c) class_start[c + 1] += class_start[c];
    int *fill = (int *)malloc(sizeof(int) * (size_t)(n_objects + 1));
    memcpy(fill, class_start, sizeof(int) * (size_t)(n_objects + 1));
    for (int i = 0; i < n; ++i)
        if (cls[i] >= 0) order[fill[cls[i]]++] = i;
    free(fill);
    return class_start[n_objects + 1];
}

int orc_features_in_area(const orc_keypoint *kps, const int *cell_start, const int *indices, float min_x, float min_y,
                         float inv_w, float inv_h, int cols, int rows, float x, float y, float r, int min_level,
                         int max_level, int cap, int *out)
{
    int n = 0;
    const int nMinCellX = (int)fmax(0.0, floor((double)((x - min_x - r) * inv_w)));
    if (nMinCellX >= cols) return 0;
    const int nMaxCellX = (int)fmin((double)(cols - 1), ceil((double)((x - min_x + r) * inv_w)));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = (int)fmax(0.0, floor((double)((y - min_y - r) * inv_h)));
    if (nMinCellY >= rows) return 0;
    const int nMaxCellY = (int)fmin((double)(rows - 1), ceil((double)((y - min_y + r) * inv_h)));
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (min_level > 0) || (max_level >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
            for (int j = cell_start[ix * rows + iy]; j < cell_start[ix * rows + iy + 1]; ++j) {
                const orc_keypoint *k = &kps[indices[j]];
                if (bCheckLevels) {
                    if (k->octave < min_level) continue;
                    if (max_level >= 0 && k->octave > max_level) continue;
                }
                const float distx = k->x - x, disty = k->y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) {
                    if (n < cap) out[n] = indices[j];
                    ++n;
                }
            }
    return n;
}

/* ------------------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
 * (ORBmatcher.cc:68-155; RadiusByViewingCos :252-258; TH_HIGH = 100 :58), SURVEY.md 8f row 3 — the local-map
 * search of Tracking::SearchLocalPoints (Tracking.cc:3354).  The MapPoint / Frame objects are flattened into the
 * plain fields the routine reads; the loop below is the reference's, statement for statement, including its
 * order dependence (a feature taken by an earlier map point with observations is skipped by later ones).
 * ------------------------------------------------------------------------------------------------------------ */
static int search_by_projection_core(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                             const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                             float inv_h, int cols, int rows, const orc_map_point *mps, const uint8_t *mp_desc,
                             int n_mp, const uint8_t *occupied, const float *scale_factors, float th, float nn_ratio,
                             int *match_of_mp, float fixed_window, int level_up, int th_dist)
{
    int nmatches = 0;
    const int bFactor = th != 1.0;
    uint8_t *blocked = (uint8_t *)malloc((size_t)(n > 0 ? n : 1)); /* F.mvpMapPoints[idx] && Observations() > 0 */
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) blocked[i] = occupied ? occupied[i] : 0;
    for (int iMP = 0; iMP < n_mp; ++iMP) {
        const orc_map_point *pMP = &mps[iMP];
        match_of_mp[iMP] = -1;
        if (!(pMP->flags & 1)) continue; /* !mbTrackInView || isBad() (:76-81) */
        const int nPredictedLevel = pMP->level;
        float r = ((double)pMP->view_cos > 0.998) ? 2.5f : 4.0f; /* RadiusByViewingCos (:252-258) */
        if (bFactor) r *= th;
        /* window: r * scale (local map, :92) or the literal 5 px of the object-point overload (:179) */
        const float rad = fixed_window > 0 ? fixed_window : r * scale_factors[nPredictedLevel];
        const int nc = orc_features_in_area(kps_un, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                                            pMP->proj_x, pMP->proj_y, rad, nPredictedLevel - 1, nPredictedLevel + level_up, n, tmp);
        if (nc == 0) continue;
        const uint8_t *MPdescriptor = mp_desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int v = 0; v < nc; ++v) {
            const int idx = tmp[v];
            if (blocked[idx]) continue; /* (:111-113) */
            if (u_right[idx] > 0) {
                const float er = fabsf(pMP->proj_xr - u_right[idx]);
                if (er > r * scale_factors[nPredictedLevel]) continue;
            }
            const int dist = orc_hamming256(MPdescriptor, frame_desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestLevel2 = bestLevel;
                bestLevel = kps_un[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = kps_un[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= th_dist) { /* TH_HIGH (:142) or TH_HIGH_FORDYNAMIC (:236) */
            if (bestLevel == bestLevel2 && (float)bestDist > nn_ratio * (float)bestDist2) continue;
            match_of_mp[iMP] = bestIdx; /* F.mvpMapPoints[bestIdx] = pMP (:148) */
            if (pMP->flags & 2) blocked[bestIdx] = 1; /* pMP->Observations() > 0: later map points skip this feature */
            nmatches++;
        }
    }
    free(blocked); free(tmp);
    return nmatches;
}

int orc_search_by_projection(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                             const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                             float inv_h, int cols, int rows, const orc_map_point *mps, const uint8_t *mp_desc,
                             int n_mp, const uint8_t *occupied, const float *scale_factors, float th, float nn_ratio,
                             int *match_of_mp)
{
    return search_by_projection_core(kps_un, frame_desc, u_right, n, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                                     mps, mp_desc, n_mp, occupied, scale_factors, th, nn_ratio, match_of_mp, 0.f, 0, 100);
}

/* ORBmatcher::SearchByProjection(Frame &F, const size_t &nOrder, const vector<MapObjectPoint*> &vpMapPoints, const float th)
 * (ORBmatcher.cc:157-246): the local-map search over ONE detected object's own feature set (mvObjKeysUn[nOrder] gridded
 * in mvObjKeysGrid[nOrder], Frame::GetObjectFeaturesInArea :1892-1951 = GetFeaturesInArea on that set) with a fixed 5 px
 * window, levels [L - 1, L + 1], TH_HIGH_FORDYNAMIC = 130; the stereo test still uses r * scale.  The caller folds
 * !isInBBox(keypoint) (:206, Frame.cc:1796-1806) into `occupied`: it is a per-feature filter. */
int orc_search_by_projection_obj(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                                 const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                                 float inv_h, int cols, int rows, const orc_map_point *mps, const uint8_t *mp_desc,
                                 int n_mp, const uint8_t *occupied, const float *scale_factors, float th, float nn_ratio,
                                 int *match_of_mp)
{
    return search_by_projection_core(kps_un, frame_desc, u_right, n, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows,
                                     mps, mp_desc, n_mp, occupied, scale_factors, th, nn_ratio, match_of_mp, 5.f, 1, 130);
}

/* ------------------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * (ORBmatcher.cc:1613-1756) from the point where each LastFrame map point has been projected: the loop body from the
 * GetFeaturesInArea call on (:1671-1727), then the rotation consistency step (:1730-1753) with ComputeThreeMaxima
 * (:2658-2699).  The projection itself (:1623-1667) stays with the caller, who passes its results per point.
 * ------------------------------------------------------------------------------------------------------------ */
int orc_search_by_projection_last(const orc_keypoint *kps_un, const uint8_t *frame_desc, const float *u_right, int n,
                                  const int *cell_start, const int *indices, float min_x, float min_y, float inv_w,
                                  float inv_h, int cols, int rows, const orc_proj_point *pts, const uint8_t *pt_desc,
                                  int n_pts, const uint8_t *occupied, int check_orientation, int *match_of_pt,
                                  uint8_t *rejected)
{
    enum { HISTO_LENGTH = 30 };
    int nmatches = 0;
    const float factor = HISTO_LENGTH / 360.0f;
    uint8_t *blocked = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *hist = (int *)malloc(sizeof(int) * (size_t)HISTO_LENGTH * (size_t)(n_pts > 0 ? n_pts : 1)); /* rotHist[bin][k] */
    int hsize[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) hsize[i] = 0;
    for (int i = 0; i < n; ++i) blocked[i] = occupied ? occupied[i] : 0;
    for (int i = 0; i < n_pts; ++i) {
        const orc_proj_point *p = &pts[i];
        match_of_pt[i] = -1;
        rejected[i] = 0;
        if (!(p->flags & 1)) continue; /* no map point, outlier, behind the camera or outside the image (:1640-1662) */
        const float radius = p->radius;
        const int nc = orc_features_in_area(kps_un, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows, p->u, p->v,
                                            radius, p->min_level, p->max_level, n, tmp);
        if (nc == 0) continue;
        const uint8_t *dMP = pt_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int v = 0; v < nc; ++v) {
            const int i2 = tmp[v];
            if (blocked[i2]) continue; /* (:1684-1686) */
            if (u_right[i2] > 0) {
                const float er = fabsf(p->ur - u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_hamming256(dMP, frame_desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= 100) { /* TH_HIGH */
            match_of_pt[i] = bestIdx2; /* CurrentFrame.mvpMapPoints[bestIdx2] = pMP (:1715) */
            if (p->flags & 2) blocked[bestIdx2] = 1;
            nmatches++;
            if (check_orientation) {
                float rot = p->angle - kps_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round((double)(rot * factor));
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin < 0 || bin >= HISTO_LENGTH) { nmatches = -1; goto done; } /* the reference asserts */
                hist[(size_t)bin * (size_t)n_pts + (size_t)hsize[bin]++] = i; /* the reference stores bestIdx2; i identifies it */
            }
        }
    }
    if (check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) { /* ComputeThreeMaxima (:2658-2699) */
            const int s = hsize[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j = 0; j < hsize[i]; j++) {
                    rejected[hist[(size_t)i * (size_t)n_pts + (size_t)j]] = 1; /* mvpMapPoints[...] = NULL (:1747) */
                    nmatches--;
                }
    }
done:
    free(blocked); free(tmp); free(hist);
    return nmatches;
}

/* ------------------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th,
 * const int ORBdist) (ORBmatcher.cc:2529-2654, relocalisation) from the point where each key-frame map point has been
 * projected (:2551-2581 stay with the caller): window search over levels [nPredictedLevel - 1, nPredictedLevel + 1]
 * (:2583), best unassigned feature (:2593-2609; ANY existing association blocks a feature, and so does every match made
 * on the way), acceptance at bestDist <= ORBdist (:2611), rotation histogram (:2616-2651).  No stereo check in this
 * overload.  pts[i]: u, v, radius = th * mvScaleFactors[level], min/max level, angle = pKF->mvKeysUn[i].angle, flags bit
 * 0 = the point reaches the search (not null / bad / already found, inside the image and the distance range).
 * ------------------------------------------------------------------------------------------------------------ */
int orc_search_by_projection_kf(const orc_keypoint *kps_un, const uint8_t *frame_desc, int n, const int *cell_start,
                                const int *indices, float min_x, float min_y, float inv_w, float inv_h, int cols, int rows,
                                const orc_proj_point *pts, const uint8_t *pt_desc, int n_pts, const uint8_t *occupied,
                                int orb_dist, int check_orientation, int *match_of_pt, uint8_t *rejected)
{
    enum { HISTO_LENGTH = 30 };
    int nmatches = 0;
    const float factor = 1.0f / HISTO_LENGTH; /* :2542 - this overload divides by 30, the last-frame one by 360/30 */
    uint8_t *blocked = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
    int *tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *hist = (int *)malloc(sizeof(int) * (size_t)HISTO_LENGTH * (size_t)(n_pts > 0 ? n_pts : 1));
    int hsize[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; ++i) hsize[i] = 0;
    for (int i = 0; i < n; ++i) blocked[i] = occupied ? occupied[i] : 0;
    for (int i = 0; i < n_pts; ++i) {
        const orc_proj_point *p = &pts[i];
        match_of_pt[i] = -1;
        rejected[i] = 0;
        if (!(p->flags & 1)) continue;
        const int nc = orc_features_in_area(kps_un, cell_start, indices, min_x, min_y, inv_w, inv_h, cols, rows, p->u, p->v,
                                            p->radius, p->min_level, p->max_level, n, tmp);
        if (nc == 0) continue;
        const uint8_t *dMP = pt_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int v = 0; v < nc; ++v) {
            const int i2 = tmp[v];
            if (blocked[i2]) continue; /* if(CurrentFrame.mvpMapPoints[i2]) continue; (:2596) */
            const int dist = orc_hamming256(dMP, frame_desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            match_of_pt[i] = bestIdx2;
            blocked[bestIdx2] = 1; /* CurrentFrame.mvpMapPoints[bestIdx2] = pMP (:2613) */
            nmatches++;
            if (check_orientation) {
                float rot = p->angle - kps_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)round((double)(rot * factor));
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin < 0 || bin >= HISTO_LENGTH) { nmatches = -1; goto done; } /* the reference asserts */
                hist[(size_t)bin * (size_t)n_pts + (size_t)hsize[bin]++] = i;
            }
        }
    }
    if (check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = hsize[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j = 0; j < hsize[i]; j++) {
                    rejected[hist[(size_t)i * (size_t)n_pts + (size_t)j]] = 1;
                    nmatches--;
                }
    }
done:
    free(blocked); free(tmp); free(hist);
    return nmatches;
}

/* ------------------------------------------------------------------------------------------------------------
 * cv::ORB::create(nfeatures, 1.2f, nlevels, edgeThreshold)->detectAndCompute(image, mask) — the second extractor of
 * PointSLOT (reference src/Frame.cc:2623-2627).  The algorithm is OpenCV's
 * (modules/features2d/src/orb.cpp: ORB_Impl::detectAndCompute, computeKeyPoints, HarrisResponses, ICAngles,
 * computeOrbDescriptors; OpenCV 4.13 semantics); this restatement is pinned against cv2.ORB_create by
 * tests/test_oracle_cvorb.py, through oracle/cvorb_path.py which established the arithmetic:
 *   - pyramid: INTER_LINEAR_EXACT, each level from the previous one (8-bit coefficients, (V + 2^15) >> 16);
 *   - blur: float separable filter with fused multiply-adds (GaussianBlur on a submatrix), not the fixed-point one;
 *   - steering: cosf / sinf of the float angle (glibc's algorithm, restated above: orc_glibc_cosf).
 * Output order: (octave, y, x) — cv::ORB's own order is what std::nth_element leaves and is not part of the contract.
 * ------------------------------------------------------------------------------------------------------------ */
static void resize_linear_exact_u8(const uint8_t *src, int sw, int sh, ptrdiff_t sp, uint8_t *dst, int dw, int dh, ptrdiff_t dp)
{
    int *x0 = malloc(sizeof(int) * (size_t)dw * 2), *y0 = malloc(sizeof(int) * (size_t)dh * 2);
    int *ax = x0 + dw, *ay = y0 + dh;
    for (int pass = 0; pass < 2; ++pass) {
        const int nd = pass ? dh : dw, ns = pass ? sh : sw;
        int *s0 = pass ? y0 : x0, *a = pass ? ay : ax;
        const double scale = (double)ns / (double)nd;
        for (int d = 0; d < nd; ++d) {
            double f = (d + 0.5) * scale - 0.5;
            int s = (int)floor(f);
            double fr = f - s;
            if (s < 0) { s = 0; fr = 0.0; }
            if (s >= ns - 1) { s = ns - 1; fr = 0.0; }
            s0[d] = s;
            a[d] = (int)lrint(fr * 256.0);
        }
    }
    int *H0 = malloc(sizeof(int) * (size_t)dw), *H1 = malloc(sizeof(int) * (size_t)dw);
    for (int y = 0; y < dh; ++y) {
        const uint8_t *r0 = src + (ptrdiff_t)y0[y] * sp, *r1 = src + (ptrdiff_t)(y0[y] + 1 < sh ? y0[y] + 1 : sh - 1) * sp;
        for (int x = 0; x < dw; ++x) {
            const int xa = x0[x], xb = xa + 1 < sw ? xa + 1 : sw - 1;
            H0[x] = r0[xa] * (256 - ax[x]) + r0[xb] * ax[x];
            H1[x] = r1[xa] * (256 - ax[x]) + r1[xb] * ax[x];
            dst[(ptrdiff_t)y * dp + x] = (uint8_t)((H0[x] * (256 - ay[y]) + H1[x] * ay[y] + (1 << 15)) >> 16);
        }
    }
    free(x0); free(y0); free(H0); free(H1);
}


/* float separable Gaussian 7x7 sigma 2 as cv::sepFilter2D evaluates it on CPUs with FMA dispatch: row taps left to
 * right, column taps from the centre outwards over the symmetric pairs, every multiply-add fused */
static void blur_float_separable(const uint8_t *src, int w, int h, ptrdiff_t sp, uint8_t *dst)
{
    /* getGaussianKernel(7, 2, CV_32F): exp(-x^2 / 8) normalised in double, stored as float */
    float k[7];
    {
        double g[7], sum = 0;
        for (int i = 0; i < 7; ++i) { const double x = i - 3; g[i] = exp(-0.125 * x * x); sum += g[i]; }
        for (int i = 0; i < 7; ++i) k[i] = (float)(g[i] / sum);
    }
    float *row = malloc(sizeof(float) * (size_t)w * (size_t)(h + 6));
    for (int y = -3; y < h + 3; ++y) {
        const uint8_t *r = src + (ptrdiff_t)reflect101(y, h) * sp;
        float *o = row + (size_t)(y + 3) * (size_t)w;
        for (int x = 0; x < w; ++x) {
            float acc = k[0] * (float)r[reflect101(x - 3, w)];
            for (int i = 1; i < 7; ++i) acc = fmaf(k[i], (float)r[reflect101(x - 3 + i, w)], acc);
            o[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float *c = row + (size_t)(y + 3) * (size_t)w + x;
            float acc = k[3] * c[0];
            for (int i = 1; i <= 3; ++i) acc = fmaf(k[3 - i], c[-(ptrdiff_t)i * w] + c[(ptrdiff_t)i * w], acc);
            long v = lrintf(acc);
            dst[(size_t)y * (size_t)w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    free(row);
}

typedef struct { float x, y, resp; } cvorb_pt;
static int cvorb_cmp_desc(const void *a, const void *b)
{
    const float x = *(const float *)a, y = *(const float *)b;
    return x > y ? -1 : x < y ? 1 : 0;
}
/* KeyPointsFilter::retainBest: keep everything whose response is >= the n-th largest (ties kept); order preserved */
static int cvorb_retain_best(cvorb_pt *p, int n, int keep)
{
    if (keep < 0 || n <= keep) return n;
    if (keep == 0) return 0;
    float *r = malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i) r[i] = p[i].resp;
    qsort(r, (size_t)n, sizeof(float), cvorb_cmp_desc);
    const float thr = r[keep - 1];
    free(r);
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (p[i].resp >= thr) p[m++] = p[i];
    return m;
}

int orc_cvorb_detect_and_compute(const uint8_t *img, int w, int h, ptrdiff_t pitch, const uint8_t *mask,
                                 ptrdiff_t mask_pitch, int nfeatures, float scale_factor_f, int nlevels,
                                 int edge_threshold, int fast_threshold, orc_keypoint *kps, uint8_t *desc32, int cap,
                                 int *n_out)
{
    if (nlevels < 1 || nlevels > MAX_LEVELS || w <= 0 || h <= 0) return -1;
    /* the same ranges as pslot_cvorb_create.  edgeThreshold < 19: the steered pattern (radius 18.4) leaves the level and
     * cv::ORB samples the unblurred reflected frame of its pyramid buffer there, which this restatement does not model */
    if (nfeatures < 1 || !(scale_factor_f > 1.0f) || edge_threshold < 19 || fast_threshold < 1 || fast_threshold > 254) return -1;
    const double scaleFactor = (double)scale_factor_f; /* ORB::create takes a float, the member is a double */
    const int patchSize = 31, halfPatch = 15;
    int umax[HALF_PATCH_SIZE + 2];
    {
        int v, v0, vmax = (int)floor(halfPatch * sqrt(2.f) / 2 + 1), vmin = (int)ceil(halfPatch * sqrt(2.f) / 2);
        for (v = 0; v <= vmax; ++v) umax[v] = orc_cv_round(sqrt((double)halfPatch * halfPatch - v * v));
        for (v = halfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    /* feature quota per level (computeKeyPoints) */
    int quota[MAX_LEVELS];
    {
        const float factor = (float)(1.0 / scaleFactor);
        float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) { quota[l] = orc_cv_round(nd); sum += quota[l]; nd *= factor; }
        quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    uint8_t *prev = NULL, *prevm = NULL;
    int pw = 0, ph = 0, n = 0, rc = 0;
    const float deg2rad = (float)(3.1415926535897932384626433832795 / 180.f);
    for (int level = 0; level < nlevels && rc == 0; ++level) {
        const float scale = (float)pow(scaleFactor, (double)level);
        /* level size = cvRound(dim * (1.f / scale)): one float reciprocal, then a float product - what OpenCV 4.13 (cv2) is
         * observed to compute; a float division differs where dim / scale lands on x.5 (465 / 1.2f -> 387,
         * 465 * (1 / 1.2f) -> 388).  Pinned against cv2.ORB_create by tools/cvorb_fuzz.py */
        const float inv_scale = 1.0f / scale;
        const int lw = orc_cv_round((float)w * inv_scale), lh = orc_cv_round((float)h * inv_scale);
        if (lw < 2 * edge_threshold + 1 || lh < 2 * edge_threshold + 1) break; /* nothing can pass the border filter */
        uint8_t *cur = malloc((size_t)lw * (size_t)lh), *curm = mask ? malloc((size_t)lw * (size_t)lh) : NULL;
        if (level == 0) {
            for (int y = 0; y < lh; ++y) {
                memcpy(cur + (size_t)y * lw, img + (ptrdiff_t)y * pitch, (size_t)lw);
                if (mask) /* cv::ORB looks at the mask as zero / non-zero: any non-zero value acts as 255 on every level */
                    for (int x = 0; x < lw; ++x) curm[(size_t)y * lw + x] = mask[(ptrdiff_t)y * mask_pitch + x] ? 255 : 0;
            }
        } else {
            resize_linear_exact_u8(prev, pw, ph, pw, cur, lw, lh, lw);
            if (mask) {
                resize_linear_exact_u8(prevm, pw, ph, pw, curm, lw, lh, lw);
                for (size_t i = 0; i < (size_t)lw * lh; ++i) curm[i] = curm[i] > 254 ? curm[i] : 0; /* THRESH_TOZERO 254 */
            }
        }
        /* FAST + mask + border filter + retainBest(2 * quota) by the FAST score */
        int cap_f = 1 << 14, nf;
        int *fx, *fy, *fr;
        for (;;) {
            fx = malloc(sizeof(int) * 3 * (size_t)cap_f); fy = fx + cap_f; fr = fy + cap_f;
            nf = orc_fast9_16(cur, lw, lh, lw, fast_threshold, cap_f, fx, fy, fr);
            if (nf <= cap_f) break;
            free(fx); cap_f = nf;
        }
        cvorb_pt *p = malloc(sizeof(cvorb_pt) * (size_t)(nf > 0 ? nf : 1));
        int m = 0;
        for (int i = 0; i < nf; ++i) {
            if (curm && curm[(size_t)fy[i] * lw + fx[i]] == 0) continue; /* runByPixelsMask */
            if (fx[i] < edge_threshold || fx[i] >= lw - edge_threshold || fy[i] < edge_threshold || fy[i] >= lh - edge_threshold)
                continue; /* runByImageBorder */
            p[m].x = (float)fx[i]; p[m].y = (float)fy[i]; p[m].resp = (float)fr[i]; ++m;
        }
        free(fx);
        m = cvorb_retain_best(p, m, 2 * quota[level]);
        /* HarrisResponses: 7x7 block, k = 0.04, float score in OpenCV's evaluation order */
        {
            const int blockSize = 7, r = blockSize / 2;
            const float harris_k = 0.04f;
            float scale1 = 1.f / ((1 << 2) * blockSize * 255.f);
            const float scale_sq_sq = scale1 * scale1 * scale1 * scale1;
            for (int i = 0; i < m; ++i) {
                const int x0 = (int)p[i].x, y0 = (int)p[i].y;
                int a = 0, b = 0, c = 0;
                for (int dy = -r; dy <= r; ++dy)
                    for (int dx = -r; dx <= r; ++dx) {
                        const uint8_t *q = cur + (size_t)(y0 + dy) * lw + (x0 + dx);
                        const int Ix = (q[1] - q[-1]) * 2 + (q[-lw + 1] - q[-lw - 1]) + (q[lw + 1] - q[lw - 1]);
                        const int Iy = (q[lw] - q[-lw]) * 2 + (q[lw - 1] - q[-lw - 1]) + (q[lw + 1] - q[-lw + 1]);
                        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
                    }
                p[i].resp = ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
            }
        }
        m = cvorb_retain_best(p, m, quota[level]);
        uint8_t *blur = malloc((size_t)lw * (size_t)lh);
        blur_float_separable(cur, lw, lh, lw, blur);
        const float inv = 1.f / scale;
        for (int i = 0; i < m && rc == 0; ++i) { /* p is still in raster order: (y, x) ascending */
            if (n >= cap) { rc = -2; break; }
            const int x0 = (int)p[i].x, y0 = (int)p[i].y;
            orc_keypoint *K = &kps[n];
            K->angle = ic_angle(cur, lw, x0, y0, umax);
            K->x = p[i].x * scale; K->y = p[i].y * scale;
            K->size = patchSize * scale; K->response = p[i].resp; K->octave = level; K->class_id = -1;
            /* computeOrbDescriptors */
            const int cy = orc_cv_round(K->y * inv), cx = orc_cv_round(K->x * inv);
            const float ang = K->angle * deg2rad;
            const float a = orc_glibc_cosf(ang), b = orc_glibc_sinf(ang);
            uint8_t *d = desc32 + (size_t)n * 32;
            for (int j = 0; j < 32; ++j) {
                int byte = 0;
                for (int t = 0; t < 8; ++t) {
                    const int8_t *pt = k_pattern + (size_t)(j * 8 + t) * 4;
                    const float ax0 = pt[0] * a - pt[1] * b, ay0 = pt[0] * b + pt[1] * a;
                    const float ax1 = pt[2] * a - pt[3] * b, ay1 = pt[2] * b + pt[3] * a;
                    const int t0 = blur[(size_t)(cy + (int)lrintf(ay0)) * lw + (cx + (int)lrintf(ax0))];
                    const int t1 = blur[(size_t)(cy + (int)lrintf(ay1)) * lw + (cx + (int)lrintf(ax1))];
                    byte |= (t0 < t1) << t;
                }
                d[j] = (uint8_t)byte;
            }
            ++n;
        }
        free(blur); free(p);
        free(prev); free(prevm);
        prev = cur; prevm = curm; pw = lw; ph = lh;
    }
    free(prev); free(prevm);
    *n_out = n;
    return rc;
}


/* ------------------------------------------------------------------------------------------------------------
 * Frame::ComputeBoW (reference src/Frame.cc:2040-2047) -> DBoW2 transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h)
 * ------------------------------------------------------------------------------------------------------------ */
void orc_bow_transform(const orc_vocab *v, const uint8_t *desc, int n, int levelsup, int32_t *leaf_node, int32_t *nid)
{
    const int nid_level = v->L - levelsup; /* :1232 */
    for (int i = 0; i < n; ++i) {
        const uint8_t *f = desc + (size_t)i * 32;
        int final_id = 0, level = 0, at = -1;
        if (nid_level <= 0) at = 0; /* :1233 */
        do { /* :1238-1261 */
            ++level;
            const int32_t *c = v->child + v->child_start[final_id], nc = v->child_start[final_id + 1] - v->child_start[final_id];
            final_id = c[0];
            int best = orc_hamming256(f, v->desc + (size_t)final_id * 32);
            for (int j = 1; j < nc; ++j) {
                const int d = orc_hamming256(f, v->desc + (size_t)c[j] * 32);
                if (d < best) { best = d; final_id = c[j]; }
            }
            if (level == nid_level) at = final_id;
        } while (v->child_start[final_id + 1] > v->child_start[final_id]);
        leaf_node[i] = final_id;
        nid[i] = at >= 0 ? at : final_id; /* the reference leaves *nid unset when the leaf lies above nid_level */
    }
}

static int cmp_i32(const void *a, const void *b) { const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return x < y ? -1 : x > y; }

int orc_bow_vectors(const orc_vocab *v, const int32_t *leaf_node, const int32_t *nid, int n, int32_t *bow_id, double *bow_val,
                    int32_t *fv_node, int32_t *fv_start, int32_t *fv_feat, int *n_fv_nodes)
{
    /* features whose word is not stopped (w > 0), :1163 / :1191 */
    int m = 0, nf = 0;
    int32_t *ids = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1)), *nodes = malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i)
        if (v->weight[leaf_node[i]] > 0) { ids[m] = v->word_id[leaf_node[i]]; nodes[m] = nid[i]; ++m; }
    /* std::map keys: the distinct ids in ascending order */
    int32_t *sorted = malloc(sizeof(int32_t) * (size_t)(m > 0 ? m : 1));
    memcpy(sorted, ids, sizeof(int32_t) * (size_t)m);
    qsort(sorted, (size_t)m, sizeof(int32_t), cmp_i32);
    int nw = 0;
    for (int i = 0; i < m; ++i)
        if (i == 0 || sorted[i] != sorted[i - 1]) { bow_id[nw] = sorted[i]; bow_val[nw] = 0.0; ++nw; }
    /* addWeight accumulates v[id] += w in feature order (TF_IDF, TF :1164); addIfNotExist stores w once (IDF, BINARY :1192) */
    const int tf = v->weighting == 0 || v->weighting == 1;
    {
        uint8_t *seen = calloc((size_t)(nw > 0 ? nw : 1), 1);
        int j = 0;
        for (int i = 0; i < n; ++i) {
            const double w = v->weight[leaf_node[i]];
            if (!(w > 0)) continue;
            const int32_t id = ids[j++];
            int lo = 0, hi = nw - 1;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (bow_id[mid] < id) lo = mid + 1; else hi = mid; }
            if (tf) { if (seen[lo]) bow_val[lo] += w; else { bow_val[lo] = w; seen[lo] = 1; } }
            else if (!seen[lo]) { bow_val[lo] = w; seen[lo] = 1; }
        }
        free(seen);
    }
    /* mustNormalize (ScoringObject.h): every scoring but DOT_PRODUCT normalises, L2_NORM by L2, the rest by L1 */
    const int must = v->scoring != 5, l2 = v->scoring == 1;
    if (tf && nw > 0 && !must) { /* :1169-1175 */
        const double nd = (double)nw;
        for (int i = 0; i < nw; ++i) bow_val[i] /= nd;
    }
    if (must) { /* BowVector::normalize (BowVector.cpp:62-85): sums in ascending id order */
        double norm = 0.0;
        if (!l2) for (int i = 0; i < nw; ++i) norm += fabs(bow_val[i]);
        else { for (int i = 0; i < nw; ++i) norm += bow_val[i] * bow_val[i]; norm = sqrt(norm); }
        if (norm > 0.0) for (int i = 0; i < nw; ++i) bow_val[i] /= norm;
    }
    /* FeatureVector (FeatureVector.cpp:33-47): node id ascending, feature indices in feature order */
    memcpy(sorted, nodes, sizeof(int32_t) * (size_t)m);
    qsort(sorted, (size_t)m, sizeof(int32_t), cmp_i32);
    for (int i = 0; i < m; ++i)
        if (i == 0 || sorted[i] != sorted[i - 1]) fv_node[nf++] = sorted[i];
    for (int i = 0; i <= nf; ++i) fv_start[i] = 0;
    for (int i = 0; i < m; ++i) {
        int lo = 0, hi = nf - 1;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (fv_node[mid] < nodes[i]) lo = mid + 1; else hi = mid; }
        ++fv_start[lo + 1];
    }
    for (int i = 0; i < nf; ++i) fv_start[i + 1] += fv_start[i];
    {
        int32_t *fill = calloc((size_t)(nf > 0 ? nf : 1), sizeof(int32_t));
        int j = 0;
        for (int i = 0; i < n; ++i) {
            if (!(v->weight[leaf_node[i]] > 0)) continue;
            const int32_t node = nodes[j++];
            int lo = 0, hi = nf - 1;
            while (lo < hi) { const int mid = (lo + hi) / 2; if (fv_node[mid] < node) lo = mid + 1; else hi = mid; }
            fv_feat[fv_start[lo] + fill[lo]++] = i;
        }
        free(fill);
    }
    free(ids); free(nodes); free(sorted);
    *n_fv_nodes = nf;
    return nw;
}
