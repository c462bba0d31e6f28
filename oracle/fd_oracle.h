/* TEST INFRASTRUCTURE -- C restatement of the reference hot path; see fd_oracle.c for the contract.
 * Signatures mirror the ref_* entry points of ref_driver.cpp one for one, so oracle/bindings.py can
 * drive either checker with the same code. */
#ifndef FD_ORACLE_H_
#define FD_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int orc_detect(int kind, const uint8_t *img, int rows, int cols, float min_response, int min_distance, uint32_t needed, int fast_n,
               float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, float *cand_resp, int32_t *cand_xy, int64_t max_cand,
               int64_t *n_cand, float *response_map, int32_t *mask_out);
void orc_fast_score_map(const uint8_t *img, int rows, int cols, int fast_n, int diff, uint8_t *score_out);
int orc_brief(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, uint8_t *bits_out);
int orc_lsd_map(const uint8_t *img, int rows, int cols, float min_norm, float *norm, float *angle, uint8_t *valid, int32_t *sorted_rc,
                int64_t max_sorted, int64_t *n_sorted);
void orc_sparsify(const float *feats_xy, int n, int rows, int cols, int grid_rows, int grid_cols, uint8_t need, uint8_t after,
                  uint8_t *status, int n_status);
int orc_brief_vec(const uint8_t *img, int rows, int cols, const float *kp_xy, int n, int length, int half_patch, float *out);
int orc_nn_select(const float *heatmap, int rows, int cols, float min_response, int invalid_boundary, int min_distance, int max_features,
                  float *feats_xy, int n_feats_in, int max_feats, int *n_feats_out, int64_t *n_candidates);
int orc_nn_descriptors(const float *feats_xy, int n_feats, const float *maps, int channels, int map_rows, int map_cols, float *out);
double orc_bench_points(int kind, const uint8_t *frames, int n_frames, int rows, int cols, float min_response, int min_distance,
                        uint32_t needed, int fast_n, int brief_length, int brief_half_patch, int n_threads, int64_t *totals);
#ifdef __cplusplus
}
#endif
#endif /* FD_ORACLE_H_ */
