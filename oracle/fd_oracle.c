/* TEST INFRASTRUCTURE -- not product code.
 *
 * Plain-C restatement of the reference's dense per-pixel hot path (Horizon1026/Feature_Detector),
 * each function citing the reference file:line it follows (paths relative to /root/reference).
 * It is the checker the CUDA path is compared with; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py may load it.  The product library never links or calls it.
 *
 * Pinning: this port is compared in tests/test_oracle.py against (a) oracle/_ref/libfd_ref.so, the
 * UNMODIFIED reference sources compiled in place, whenever that build is present, and (b) the
 * committed golden vectors under tests/golden/ that were generated from that build
 * (tests/golden/make_golden.py).  Detector, response-map and LSD-field parity is therefore pinned
 * to the reference's code.  BRIEF bits are pinned only to the reference compiled against
 * compat/slam_utility/datatype_image.h, whose float-coordinate pixel fetch is GUESS G1
 * (SURVEY.md 8c): against the real upstream Slam_Utility the BRIEF parity is UNPINNED.
 *
 * Deliberate, documented deviations from the reference binary:
 *   - ties in the candidate sort are broken in raster order (row, then col); the reference uses an
 *     unstable std::sort (feature_point_detector.cpp:58), so tied candidates may come out in another
 *     order there.  Same for the LSD seed sort (column-major push order kept; feature_line_detector.cpp:92).
 *   - structure-tensor sums are taken directly over the 3x3 window in integers instead of the
 *     reference's float sliding sums; fact H1 (every partial sum is an exact integer < 2^24) makes the
 *     two bitwise identical, and the test against _ref checks it.
 *   - a bilinear tap that would fall past the end of the image buffer (reachable only for a keypoint
 *     exactly on the border limit) reads as the last byte instead of invoking undefined behaviour.
 *
 * Compile with -ffp-contract=off: the reference build has no FMA (CMakeLists.txt:6 has no -march).
 */
#include "fd_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { ORC_HARRIS = 0, ORC_SHI_TOMAS = 1, ORC_FAST = 2 };

typedef struct {
    float resp;
    int32_t x, y;
} orc_cand;

/* ------------------------------------------------------------------------------------------------
 * mask handling -- feature_point_detector.cpp:76-98 (DrawRectangleInMask, UpdateMaskByFeatures)
 * ---------------------------------------------------------------------------------------------- */
static void clear_square(int32_t *mask, int rows, int cols, int row, int col, int d) {
    for (int dr = -d; dr <= d; ++dr) {
        for (int dc = -d; dc <= d; ++dc) {
            const int r = row + dr, c = col + dc;
            if (r < 0 || c < 0 || r > rows - 1 || c > cols - 1) continue;
            mask[(size_t)r * cols + c] = 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Harris / Shi-Tomasi -- feature_point_harris_detector.cpp:17-137, feature_point_shi_tomas_detector.cpp:94-101
 * ---------------------------------------------------------------------------------------------- */
static void corner_response_map(int kind, const uint8_t *img, int rows, int cols, const int32_t *mask, float thr, float *resp) {
    memset(resp, 0, sizeof(float) * (size_t)rows * cols);              /* harris.cpp:74-75 */
    if (rows < 5 || cols < 5) return;
    const float inv_cnt = 1.0f / 9.0f;                                 /* harris.cpp:71, kHalfPatchSize = 1 (harris.h:14) */
    const float inv_cnt2 = inv_cnt * inv_cnt;                          /* harris.cpp:72 */
    const float alpha = 0.04f;                                         /* harris.h:13 */
    for (int r = 2; r < rows - 2; ++r) {                               /* bound = half_size + 1 = 2, harris.cpp:90-92 */
        for (int c = 2; c < cols - 2; ++c) {
            if (!mask[(size_t)r * cols + c]) continue;                 /* harris.cpp:94 */
            int32_t ixx = 0, iyy = 0, ixy = 0;
            for (int dr = -1; dr <= 1; ++dr) {
                for (int dc = -1; dc <= 1; ++dc) {
                    const uint8_t *p = img + (size_t)(r + dr) * cols + (c + dc);
                    const int32_t ix = (int32_t)p[1] - (int32_t)p[-1];        /* harris.cpp:36 */
                    const int32_t iy = (int32_t)p[cols] - (int32_t)p[-cols];  /* harris.cpp:37 */
                    ixx += ix * ix;                                           /* harris.cpp:38-40 and the two sliding sums */
                    iyy += iy * iy;
                    ixy += ix * iy;
                }
            }
            const float sxx = (float)ixx, syy = (float)iyy, sxy = (float)ixy; /* exact: |sum| <= 585225 < 2^24 (H1) */
            float res;
            if (kind == ORC_HARRIS) {
                const float trace = sxx + syy;                                         /* harris.cpp:97 */
                if (!(trace * trace * 0.21f * inv_cnt2 > thr)) continue;               /* harris.cpp:98 */
                res = (sxx * syy - sxy * sxy - alpha * trace * trace) * inv_cnt2;      /* harris.cpp:100 */
            } else {
                const float a = sxx * inv_cnt;                                         /* shi_tomas.cpp:94 */
                const float c_val = syy * inv_cnt;                                     /* shi_tomas.cpp:95 */
                if (!(a + c_val > thr)) continue;                                      /* shi_tomas.cpp:96 */
                const float b = sxy * inv_cnt;                                         /* shi_tomas.cpp:97 */
                const float diff = a - c_val;                                          /* shi_tomas.cpp:98 */
                const float common = sqrtf(diff * diff + 4.0f * b * b);                /* shi_tomas.cpp:99 */
                res = (a + c_val + common) * 0.5f;                                     /* shi_tomas.cpp:100 (the LARGER eigenvalue) */
            }
            if (res > thr) resp[(size_t)r * cols + c] = res;                           /* harris.cpp:101-103 */
        }
    }
}

/* harris.cpp:120-137: strict 4-neighbour maximum on the thresholded map, raster order. */
static int64_t corner_nms(const float *resp, int rows, int cols, float thr, orc_cand *out) {
    int64_t n = 0;
    for (int r = 2; r < rows - 2; ++r) {
        const float *row = resp + (size_t)r * cols;
        for (int c = 2; c < cols - 2; ++c) {
            const float v = row[c];
            if (v <= thr) continue;                                                    /* harris.cpp:130 */
            if (v > row[c - 1] && v > row[c + 1] && v > row[c - cols] && v > row[c + cols]) { /* harris.cpp:131-132 */
                out[n].resp = v;
                out[n].x = c;
                out[n].y = r;
                ++n;
            }
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * FAST -- feature_point_fast_detector.cpp:7-98
 * ---------------------------------------------------------------------------------------------- */
static const int kRing[16][2] = {  /* {dx, dy}: index 0 = top, clockwise (fast.cpp:7-8) */
    {0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0}, {3, 1}, {2, 2}, {1, 3}, {0, 3}, {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

static int fast_score(const uint8_t *img, int cols, int row, int col, int fast_n, int diff) {
    const int32_t p = img[(size_t)row * cols + col];
    const int32_t hi = p + diff, lo = p - diff;                        /* fast.cpp:12-14: int32, no saturation */
    int larger = 0, smaller = 0;
    if (fast_n >= 12) {                                                /* fast.cpp:20-42 */
        static const int idx[4] = {0, 4, 8, 12};
        for (int i = 0; i < 4; ++i) {
            const int32_t v = img[(size_t)(row + kRing[idx[i]][1]) * cols + (col + kRing[idx[i]][0])];
            if (v > hi) { ++larger; smaller = 0; }
            else if (v < lo) { ++smaller; larger = 0; }
            else { smaller = 0; larger = 0; }
        }
        if (smaller < 3 && larger < 3) return 0;                       /* only the TRAILING run counts (fact F2) */
    }
    int cmp[16];
    for (int i = 0; i < 16; ++i) {                                     /* fast.cpp:44-52 */
        const int32_t v = img[(size_t)(row + kRing[i][1]) * cols + (col + kRing[i][0])];
        cmp[i] = (v > hi) ? 1 : ((v < lo) ? -1 : 0);
    }
    larger = smaller = 0;
    int best = 0;
    for (int k = 0; k < 2 && best < 16; ++k) {                         /* fast.cpp:55-78: walk the ring twice */
        for (int i = 0; i < 16; ++i) {
            if (cmp[i] == 1) { ++larger; smaller = 0; }
            else if (cmp[i] == -1) { ++smaller; larger = 0; }
            else { smaller = 0; larger = 0; }
            if (larger > best) best = larger;
            if (smaller > best) best = smaller;
        }
    }
    return best;                                                       /* no comparison with kN (fact F1) */
}

static int64_t fast_candidates(const uint8_t *img, int rows, int cols, const int32_t *mask, float thr, int fast_n, int diff, orc_cand *out) {
    int64_t n = 0;
    float offset = 1e-5f;                                              /* fast.cpp:85 */
    for (int row = 3; row < rows - 3; ++row) {
        for (int col = 3; col < cols - 3; ++col) {
            if (!mask[(size_t)row * cols + col]) continue;             /* fast.cpp:88 */
            const float response = (float)fast_score(img, cols, row, col, fast_n, diff) + offset; /* fast.cpp:89 */
            if (response > thr) {                                      /* fast.cpp:90-92 */
                out[n].resp = response;
                out[n].x = col;
                out[n].y = row;
                ++n;
            }
            offset += 1e-5f;                                           /* fast.cpp:93: every masked-in pixel */
        }
    }
    return n;
}

void orc_fast_score_map(const uint8_t *img, int rows, int cols, int fast_n, int diff, uint8_t *score_out) {
    if (fast_n <= 0) fast_n = 12;
    if (diff < 0) diff = 15;
    memset(score_out, 0, (size_t)rows * cols);
    for (int r = 3; r < rows - 3; ++r)
        for (int c = 3; c < cols - 3; ++c) score_out[(size_t)r * cols + c] = (uint8_t)fast_score(img, cols, r, c, fast_n, diff);
}

/* ------------------------------------------------------------------------------------------------
 * selection -- feature_point_detector.cpp:54-74
 * ---------------------------------------------------------------------------------------------- */
static void merge_sort_desc(orc_cand *a, orc_cand *tmp, int64_t n) {   /* stable: ties keep raster order */
    if (n < 2) return;
    const int64_t h = n / 2;
    merge_sort_desc(a, tmp, h);
    merge_sort_desc(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].resp > a[i].resp) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(orc_cand) * (size_t)n);
}

typedef struct {
    int32_t *mask;
    float *resp;
    orc_cand *cand, *tmp;
    size_t px;
} orc_scratch;

static int scratch_reserve(orc_scratch *s, size_t px) {
    if (px <= s->px) return 1;
    free(s->mask); free(s->resp); free(s->cand); free(s->tmp);
    s->mask = (int32_t *)malloc(sizeof(int32_t) * px);
    s->resp = (float *)malloc(sizeof(float) * px);
    s->cand = (orc_cand *)malloc(sizeof(orc_cand) * px);
    s->tmp = (orc_cand *)malloc(sizeof(orc_cand) * px);
    s->px = px;
    return s->mask && s->resp && s->cand && s->tmp;
}

static void scratch_free(orc_scratch *s) {
    free(s->mask); free(s->resp); free(s->cand); free(s->tmp);
    memset(s, 0, sizeof(*s));
}

/* FeaturePointDetector::DetectGoodFeatures, feature_point_detector.cpp:7-25.  Returns the number of
 * candidates; features are appended to feats_xy (n_feats in/out).  -1 when feats_xy is too small. */
static int64_t detect_core(orc_scratch *s, int kind, const uint8_t *img, int rows, int cols, float thr, int min_distance, uint32_t needed,
                           int fast_n, float *feats_xy, int *n_feats, int max_feats) {
    const size_t px = (size_t)rows * cols;
    if (!scratch_reserve(s, px)) return -3;
    for (size_t i = 0; i < px; ++i) s->mask[i] = 1;                    /* :12-13 / :91 */
    for (int i = 0; i < *n_feats; ++i) {                               /* :93-97, float -> int truncation */
        const int row = (int)feats_xy[2 * i + 1];
        const int col = (int)feats_xy[2 * i];
        clear_square(s->mask, rows, cols, row, col, min_distance);
    }
    int64_t n;
    if (kind == ORC_FAST) {
        n = fast_candidates(img, rows, cols, s->mask, thr, fast_n > 0 ? fast_n : 12, 15, s->cand);
    } else {
        corner_response_map(kind, img, rows, cols, s->mask, thr, s->resp);
        n = corner_nms(s->resp, rows, cols, thr, s->cand);
    }
    if (n == 0) return 0;                                              /* :55 */
    merge_sort_desc(s->cand, s->tmp, n);                               /* :58-60 */
    for (int64_t i = 0; i < n; ++i) {                                  /* :62-71 */
        const int row = s->cand[i].y, col = s->cand[i].x;
        if (!s->mask[(size_t)row * cols + col]) continue;
        if (*n_feats >= max_feats) return -1;
        feats_xy[2 * *n_feats] = (float)col;                           /* Vec2(pixel.x(), pixel.y()), :67 */
        feats_xy[2 * *n_feats + 1] = (float)row;
        ++*n_feats;
        if ((uint32_t)*n_feats >= needed) break;                       /* :68: tested AFTER the push */
        clear_square(s->mask, rows, cols, row, col, min_distance);     /* :69 */
    }
    return n;
}

int orc_detect(int kind, const uint8_t *img, int rows, int cols, float min_response, int min_distance, uint32_t needed, int fast_n,
               float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, float *cand_resp, int32_t *cand_xy, int64_t max_cand,
               int64_t *n_cand, float *response_map, int32_t *mask_out) {
    if (img == NULL) return 0;                                         /* :9 */
    orc_scratch s;
    memset(&s, 0, sizeof(s));
    int nf = n_feats_in;
    const int64_t n = detect_core(&s, kind, img, rows, cols, min_response, min_distance, needed, fast_n, feats_xy, &nf, max_feats);
    int rc = 1;
    if (n < 0) {
        rc = (int)n;
    } else {
        *n_feats_out = nf;
        if (n_cand) *n_cand = n;
        if (cand_resp && cand_xy) {
            if (n > max_cand) {
                rc = -2;
            } else {
                for (int64_t i = 0; i < n; ++i) {
                    cand_resp[i] = s.cand[i].resp;
                    cand_xy[2 * i] = s.cand[i].x;
                    cand_xy[2 * i + 1] = s.cand[i].y;
                }
            }
        }
        if (response_map) {
            if (kind == ORC_FAST) memset(response_map, 0, sizeof(float) * (size_t)rows * cols);
            else memcpy(response_map, s.resp, sizeof(float) * (size_t)rows * cols);
        }
        if (mask_out) memcpy(mask_out, s.mask, sizeof(int32_t) * (size_t)rows * cols);
    }
    scratch_free(&s);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * SparsifyFeatures -- feature_point_detector.cpp:27-52
 * ---------------------------------------------------------------------------------------------- */
void orc_sparsify(const float *feats_xy, int n, int rows, int cols, int grid_rows, int grid_cols, uint8_t need, uint8_t after,
                  uint8_t *status, int n_status) {
    (void)n_status;
    const float row_step = (float)(rows / (grid_rows - 1));            /* :34 integer division, then float */
    const float col_step = (float)(cols / (grid_cols - 1));            /* :35 */
    uint8_t *grid = (uint8_t *)malloc((size_t)grid_rows * grid_cols);
    memset(grid, 1, (size_t)grid_rows * grid_cols);                    /* :36 */
    for (int i = 0; i < n; ++i) {
        const int row = (int)(feats_xy[2 * i + 1] / row_step);         /* :38 */
        const int col = (int)(feats_xy[2 * i] / col_step);             /* :39 */
        if (row < 0 || row > grid_rows - 1 || col < 0 || col > grid_cols - 1) {  /* :41-44 */
            status[i] = after;
            continue;
        }
        uint8_t *cell = grid + (size_t)row * grid_cols + col;
        if (*cell && status[i] == need) *cell = 0;                     /* :46-47 */
        else if (!*cell && status[i] == need) status[i] = after;       /* :48-49 */
    }
    free(grid);
}

/* ------------------------------------------------------------------------------------------------
 * BRIEF -- descriptor_brief.cpp:8-50, descriptor.h:28-40
 * ---------------------------------------------------------------------------------------------- */
static const int8_t kPattern[256][4] = {
#include "brief_pattern_256.inc"
};

/* compat/slam_utility/datatype_image.h GetPixelValueNoCheck(float row, float col) -- GUESS G1. */
static float sample_bilinear(const uint8_t *img, int rows, int cols, float row, float col) {
    const int64_t last = (int64_t)rows * cols - 1;
    int64_t base = (int64_t)(int32_t)row * cols + (int32_t)col;
    int64_t i0 = base, i1 = base + 1, i2 = base + cols, i3 = base + cols + 1;
    if (i0 < 0) i0 = 0;
    if (i1 < 0) i1 = 0;
    if (i2 < 0) i2 = 0;
    if (i3 < 0) i3 = 0;
    if (i0 > last) i0 = last;
    if (i1 > last) i1 = last;
    if (i2 > last) i2 = last;
    if (i3 > last) i3 = last;
    const float sub_row = row - floorf(row);
    const float sub_col = col - floorf(col);
    const float inv_sub_row = 1.0f - sub_row;
    const float inv_sub_col = 1.0f - sub_col;
    return inv_sub_col * inv_sub_row * (float)img[i0] + sub_col * inv_sub_row * (float)img[i1] + inv_sub_col * sub_row * (float)img[i2] +
           sub_col * sub_row * (float)img[i3];
}

static void brief_one(const uint8_t *img, int rows, int cols, float u, float v, int length, int half_patch, uint8_t *bits) {
    memset(bits, 0, (size_t)length);                                   /* brief.cpp:10 */
    const float max_bound = fmaxf(19.0f, (float)half_patch * 2.0f);    /* brief.cpp:13-14 */
    if (u < max_bound || u > (float)cols - max_bound || v < max_bound || v > (float)rows - max_bound) return; /* :15-17 */
    float m01 = 0.0f, m10 = 0.0f;
    for (int dx = -half_patch; dx <= half_patch; ++dx) {               /* brief.cpp:22-28, dx outer */
        for (int dy = -half_patch; dy <= half_patch; ++dy) {
            const float value = sample_bilinear(img, rows, cols, v + (float)dy, u + (float)dx);
            m10 += (float)dx * value;
            m01 += (float)dy * value;
        }
    }
    const float m = sqrtf(m01 * m01 + m10 * m10);                      /* brief.cpp:29 */
    if (m < 1e-6f) return;                                             /* brief.cpp:30, kZeroFloat (G2) */
    const float sin_theta = m01 / m;                                   /* brief.cpp:32 */
    const float cos_theta = m10 / m;                                   /* brief.cpp:33 */
    const float neg_sin = -sin_theta;
    for (int i = 0; i < length; ++i) {                                 /* brief.cpp:38-47 */
        const float px1 = (float)kPattern[i][0], py1 = (float)kPattern[i][1];
        const float px2 = (float)kPattern[i][2], py2 = (float)kPattern[i][3];
        const float x1 = (cos_theta * px1 + neg_sin * py1) + u;        /* rot * Vec2 + pixel_uv, brief.cpp:35,40 */
        const float y1 = (sin_theta * px1 + cos_theta * py1) + v;
        const float x2 = (cos_theta * px2 + neg_sin * py2) + u;        /* brief.cpp:41 */
        const float y2 = (sin_theta * px2 + cos_theta * py2) + v;
        const float value_1 = sample_bilinear(img, rows, cols, y1, x1);/* brief.cpp:42 (row = y, col = x) */
        const float value_2 = sample_bilinear(img, rows, cols, y2, x2);/* brief.cpp:43 */
        if (value_1 < value_2) bits[i] = 1;                            /* brief.cpp:44-46 */
    }
}

int orc_brief(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, uint8_t *bits_out) {
    if (n <= 0 || img == NULL) return 0;                               /* descriptor.h:29 */
    if (length > 256) length = 256;
    for (int i = 0; i < n; ++i) brief_one(img, rows, cols, kp_xy[2 * i], kp_xy[2 * i + 1], length, half_patch, bits_out + (size_t)i * length);
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * LSD gradient / level-line field -- feature_line_detector.cpp:56-97
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float norm;
    int32_t row, col;
} orc_seed;

static void seed_sort_desc(orc_seed *a, orc_seed *tmp, int64_t n) {    /* stable: ties keep the column-major push order */
    if (n < 2) return;
    const int64_t h = n / 2;
    seed_sort_desc(a, tmp, h);
    seed_sort_desc(a + h, tmp, n - h);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].norm > a[i].norm) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(orc_seed) * (size_t)n);
}

int orc_lsd_map(const uint8_t *img, int rows, int cols, float min_norm, float *norm, float *angle, uint8_t *valid, int32_t *sorted_rc,
                int64_t max_sorted, int64_t *n_sorted) {
    if (img == NULL || rows < 2 || cols < 2) return 0;                 /* .cpp:14 */
    const int mr = rows - 1, mc = cols - 1;                            /* .cpp:58 */
    memset(norm, 0, sizeof(float) * (size_t)mr * mc);
    memset(angle, 0, sizeof(float) * (size_t)mr * mc);
    memset(valid, 0, (size_t)mr * mc);
    orc_seed *seeds = (orc_seed *)malloc(sizeof(orc_seed) * ((size_t)mr * mc + 1));
    orc_seed *tmp = (orc_seed *)malloc(sizeof(orc_seed) * ((size_t)mr * mc + 1));
    int64_t n = 0;
    for (int col = 1; col < cols - 2; ++col) {                         /* .cpp:71: column outer */
        for (int row = 1; row < rows - 2; ++row) {                     /* .cpp:72 */
            const int32_t a = img[(size_t)row * cols + col], b = img[(size_t)row * cols + col + 1];
            const int32_t c = img[(size_t)(row + 1) * cols + col], d = img[(size_t)(row + 1) * cols + col + 1];
            const int32_t ad = d - a;                                  /* .cpp:76-77 */
            const int32_t bc = b - c;                                  /* .cpp:78-79 */
            const float gx = (float)(ad + bc) / 2.0f;                  /* .cpp:80 */
            const float gy = (float)(ad - bc) / 2.0f;                  /* .cpp:81 */
            const float g = sqrtf(gx * gx + gy * gy);                  /* .cpp:82 */
            const size_t i = (size_t)row * mc + col;
            norm[i] = g;
            if (g > min_norm) {                                        /* .cpp:83, strict (fact L3) */
                valid[i] = 1;
                angle[i] = atan2f(gx, -gy);                            /* .cpp:85 */
                seeds[n].norm = g;
                seeds[n].row = row;
                seeds[n].col = col;
                ++n;
            }
        }
    }
    seed_sort_desc(seeds, tmp, n);                                     /* .cpp:92-94 */
    *n_sorted = n;
    int rc = 1;
    if (sorted_rc) {
        if (n > max_sorted) rc = -1;
        else
            for (int64_t i = 0; i < n; ++i) {
                sorted_rc[2 * i] = seeds[i].row;
                sorted_rc[2 * i + 1] = seeds[i].col;
            }
    }
    free(seeds);
    free(tmp);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * CPU timing of the port (bench.py cpu_baseline kind "port"): one scratch set per thread.
 * ---------------------------------------------------------------------------------------------- */
/* Descriptor<T>::Compute, std::vector<Vec> overload (reference src/feature_descriptor/descriptor.h:43-62): +1 / -1 per bit. */
int orc_brief_vec(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, float *out) {
    uint8_t *bits = (uint8_t *)malloc((size_t)(n > 0 ? n : 1) * length);
    if (!bits) return 0;
    const int ok = orc_brief(img, rows, cols, kp_xy, n, length, half_patch, bits);
    if (ok)
        for (size_t i = 0; i < (size_t)n * length; ++i) out[i] = bits[i] ? 1.0f : -1.0f;
    free(bits);
    return ok;
}

/* ---- NN detector post-processing (reference src/nn_feature_point_detector/nn_feature_point_detector.cpp) ------------------
 * orc_nn_select: CreateMask (:59-72, with UpdateMaskByFeatures :85-91 and DrawRectangleInMask :74-83),
 * SelectKeypointCandidatesFromHeatMap (:128-139) and SelectGoodFeaturesFromCandidates (:141-155).  The reference keeps
 * the candidates in a std::multimap<float, Pixel> and walks it backwards: response descending, and among equal responses
 * the LATER insertion (raster order) first. */
typedef struct {
    float response;
    int32_t index; /* raster index: row * cols + col */
} nn_cand;

static int nn_cand_before(const void *a, const void *b) {
    const nn_cand *x = (const nn_cand *)a, *y = (const nn_cand *)b;
    if (x->response != y->response) return x->response > y->response ? -1 : 1;
    return x->index > y->index ? -1 : (x->index < y->index ? 1 : 0);
}

static void nn_clear_square(uint8_t *mask, int rows, int cols, int row, int col, int radius) { /* :74-83 */
    const int r0 = row - radius > 0 ? row - radius : 0, r1 = row + radius < rows - 1 ? row + radius : rows - 1;
    const int c0 = col - radius > 0 ? col - radius : 0, c1 = col + radius < cols - 1 ? col + radius : cols - 1;
    for (int r = r0; r <= r1; ++r)
        for (int c = c0; c <= c1; ++c) mask[(size_t)r * cols + c] = 0;
}

int orc_nn_select(const float *heatmap, int rows, int cols, float min_response, int invalid_boundary, int min_distance, int max_features,
                  float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, int64_t *n_candidates) {
    uint8_t *mask = (uint8_t *)malloc((size_t)rows * cols);
    nn_cand *cand = (nn_cand *)malloc(sizeof(nn_cand) * (size_t)rows * cols);
    if (!mask || !cand) {
        free(mask);
        free(cand);
        return 0;
    }
    memset(mask, 1, (size_t)rows * cols);
    if (invalid_boundary) { /* :62-67 */
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c)
                if (r < invalid_boundary || r >= rows - invalid_boundary || c < invalid_boundary || c >= cols - invalid_boundary) mask[(size_t)r * cols + c] = 0;
    }
    for (int i = 0; i < n_feats_in; ++i) nn_clear_square(mask, rows, cols, (int32_t)feats_xy[2 * i + 1], (int32_t)feats_xy[2 * i], min_distance); /* :85-91 */
    int64_t n = 0;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            if (heatmap[(size_t)r * cols + c] > min_response) { /* :133 */
                cand[n].response = heatmap[(size_t)r * cols + c];
                cand[n].index = r * cols + c;
                ++n;
            }
    if (n_candidates) *n_candidates = n;
    qsort(cand, (size_t)n, sizeof(nn_cand), nn_cand_before);
    int nf = n_feats_in;
    for (int64_t i = 0; i < n; ++i) { /* :145-153 */
        const int row = cand[i].index / cols, col = cand[i].index % cols;
        if (!mask[(size_t)row * cols + col]) continue;
        if (nf < max_feats) {
            feats_xy[2 * nf] = (float)col;
            feats_xy[2 * nf + 1] = (float)row;
        }
        ++nf;
        if (nf >= max_features) break; /* tested after the push, before the square is cleared */
        nn_clear_square(mask, rows, cols, row, col, min_distance);
    }
    *n_feats_out = nf < max_feats ? nf : max_feats;
    free(mask);
    free(cand);
    return 1;
}

/* ExtractDescriptorsForSelectedFeatures (:163-193): bilinear taps of every channel plane at (y / 8, x / 8); a keypoint whose
 * 2x2 neighbourhood leaves the plane gets 0 for that channel.  Products and sums in the reference's order, no contraction. */
int orc_nn_descriptors(const float *feats_xy, int n_feats, const float *maps, int channels, int map_rows, int map_cols, float *out) {
    for (int i = 0; i < n_feats; ++i) {
        const float row = feats_xy[2 * i + 1] / 8.0f, col = feats_xy[2 * i] / 8.0f;
        const int32_t int_row = (int32_t)row, int_col = (int32_t)col;
        const float sub_row = row - floorf(row), sub_col = col - floorf(col);
        const float inv_sub_row = 1.0f - sub_row, inv_sub_col = 1.0f - sub_col;
        const float w0 = inv_sub_col * inv_sub_row, w1 = sub_col * inv_sub_row, w2 = inv_sub_col * sub_row, w3 = sub_col * sub_row;
        for (int j = 0; j < channels; ++j) {
            float v = 0.0f;
            if (!(int_row < 0 || int_row >= map_rows - 1 || int_col < 0 || int_col >= map_cols - 1)) {
                const float *p = maps + (size_t)j * map_rows * map_cols + (size_t)int_row * map_cols + int_col;
                v = w0 * p[0] + w1 * p[1] + w2 * p[map_cols] + w3 * p[map_cols + 1];
            }
            out[(size_t)i * channels + j] = v;
        }
    }
    return 1;
}

typedef struct {
    int kind, n_frames, rows, cols, min_distance, fast_n, brief_length, brief_half;
    uint32_t needed;
    float thr;
    const uint8_t *frames;
    volatile int *next;
    int64_t kp, cand, ones;
} orc_job;

static void *bench_worker(void *arg) {
    orc_job *j = (orc_job *)arg;
    orc_scratch s;
    memset(&s, 0, sizeof(s));
    const int max_feats = (int)j->needed + 8;
    float *feats = (float *)malloc(sizeof(float) * 2 * (size_t)max_feats);
    uint8_t *bits = (uint8_t *)malloc((size_t)max_feats * 256);
    for (;;) {
        const int f = __sync_fetch_and_add(j->next, 1);
        if (f >= j->n_frames) break;
        const uint8_t *img = j->frames + (size_t)f * j->rows * j->cols;
        int nf = 0;
        const int64_t n = detect_core(&s, j->kind, img, j->rows, j->cols, j->thr, j->min_distance, j->needed, j->fast_n, feats, &nf, max_feats);
        j->kp += nf;
        j->cand += n > 0 ? n : 0;
        if (j->brief_length > 0 && nf > 0) {
            orc_brief(img, j->rows, j->cols, feats, nf, j->brief_length, j->brief_half, bits);
            for (size_t i = 0; i < (size_t)nf * j->brief_length; ++i) j->ones += bits[i];
        }
    }
    free(feats);
    free(bits);
    scratch_free(&s);
    return NULL;
}

double orc_bench_points(int kind, const uint8_t *frames, int n_frames, int rows, int cols, float min_response, int min_distance,
                        uint32_t needed, int fast_n, int brief_length, int brief_half_patch, int n_threads, int64_t *totals) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    volatile int next = 0;
    orc_job jobs[256];
    pthread_t th[256];
    struct timespec t0, t1;
    for (int t = 0; t < n_threads; ++t) {
        orc_job j = {kind, n_frames, rows, cols, min_distance, fast_n, brief_length, brief_half_patch, needed, min_response, frames, &next, 0, 0, 0};
        jobs[t] = j;
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (n_threads == 1) {
        bench_worker(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, bench_worker, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (totals) {
        totals[0] = totals[1] = totals[2] = 0;
        for (int t = 0; t < n_threads; ++t) {
            totals[0] += jobs[t].kp;
            totals[1] += jobs[t].cand;
            totals[2] += jobs[t].ones;
        }
    }
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
