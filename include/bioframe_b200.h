/* bioframe_b200 -- C ABI of the B200-native interval engine.
 *
 * Drop-in boundary for the hot path of open2c/bioframe v0.8.0.  The reference
 * is pure Python and has no FFI of its own; its seams are the frame-level
 * helpers `_overlap_intidxs` (src/bioframe/ops.py:242-338) and
 * `_closest_intidxs` (ops.py:919-1040), the group loops of cluster/merge
 * (ops.py:637-674, 776-810), complement (ops.py:1644-1685), coverage
 * (ops.py:842-916) and the array functions of src/bioframe/core/arrops.py.
 * Each entry point below names the reference lines it replaces.  A maintainer
 * binds them with ctypes; `bioframe_b200/_lib.py` is that
 * binding.
 *
 * Conventions
 *  - Every data pointer is a DEVICE pointer owned by the caller (e.g. a torch
 *    CUDA tensor): int32 group codes, int64 starts/ends, int64 outputs.  The
 *    callee never frees or retains them.  `ws` is caller-owned device scratch of
 *    at least `*_workspace_bytes(...)` bytes; it carries state from a `_count`
 *    call to the matching `_emit` call and must not be touched in between.
 *  - Group codes are the RANK of the row's group in the order the reference
 *    would visit the groups (set order for overlap/closest, sorted key order
 *    for cluster/merge; computed by the Python glue), in [0, n_groups).  Rows
 *    of different groups never interact, so one call covers all chromosomes.
 *  - Variable-length results are two-phase: `_count` sizes the output and
 *    returns the row count (it synchronises `stream` once or twice to read it
 *    back), the caller allocates, `_emit` fills.
 *  - Return value: 0 = ok, negative = error (BF_ERR_*); `bf_last_error()` gives
 *    a thread-local message.  No C++ exception crosses the ABI.
 *  - Calls enqueue on `stream` (a cudaStream_t passed as void*), are re-entrant
 *    per (device, stream), and keep no global mutable state.
 *  - Row counts per set must be < 2^31.
 */
#ifndef BIOFRAME_B200_H
#define BIOFRAME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_ERR_ARG -1
#define BF_ERR_WORKSPACE -2
#define BF_ERR_RANGE -3
#define BF_ERR_CUDA -4
#define BF_ERR_GROUP -5
#define BF_ERR_STATE -6

/* how= of bioframe.overlap (ops.py:361) */
#define BF_HOW_INNER 0
#define BF_HOW_LEFT 1
#define BF_HOW_RIGHT 2
#define BF_HOW_OUTER 3

/* Opaque per-call state written by a *_count call and read by *_emit. */
typedef struct bf_plan {
  int64_t opaque[96];
} bf_plan;

const char* bf_last_error(void);
int bf_version(void);
/* 1 for the range-checked build (-DBF_CHECKED: device-side index checks that trap), 0 for the product build */
int bf_checked_build(void);

/* ---- launch counter and CUDA-event profiler (measurement only; used by
 * bench.py for `gpu_launches` and the per-kernel roofline).  Launches are
 * always counted per kernel class; while enabled every launch is bracketed by
 * two cudaEventRecord on its stream.  bf_prof_read synchronises the pending
 * events.  `bytes` = algorithmic bytes of the launches. */
int bf_prof_enable(int on);
int bf_prof_reset(void);
int bf_prof_num_classes(void);
const char* bf_prof_class_name(int cls);
int bf_prof_read(int cls, double* ms_total, int64_t* timed_launches, int64_t* launches, int64_t* bytes);

/* ---- overlap: replaces ops._overlap_intidxs (ops.py:242-338), i.e. the
 * per-group loop around arrops.overlap_intervals (arrops.py:290-375) plus
 * ops._minus (ops.py:228-239).  Output rows, per group in code order:
 * A-block pairs, B-block pairs (arrops.py:357-368), then (unpaired1, -1)
 * ascending (how left/outer), then (-1, unpaired2) ascending (how right/outer).
 * `closed` is the closed= flag of arrops.overlap_intervals. */
size_t bf_overlap_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups, int32_t how);
int bf_overlap_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                     const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                     int32_t n_groups, int32_t how, int32_t closed, void* ws, size_t ws_bytes,
                     bf_plan* plan, void* stream, int64_t* n_rows_out, int64_t* n_pairs_out);
int bf_overlap_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out,
                    void* stream);
/* Optional, between _count and _emit: every reported row number of set 1 / set 2 gets this offset
 * (-1 stays -1).  A query-sharded caller passes the position of its shard in the whole query set, so
 * that the concatenated per-shard results carry global row numbers without another
 * pass over the output. */
int bf_overlap_set_row_offsets(bf_plan* plan, int64_t offset1, int64_t offset2);

/* ---- prepared sets: the query-sharded, multi-GPU form of ops._overlap_intidxs (the group loop
 * of ops.py:301-338 split over ranks).  The reference sorts both sets inside every call (np.lexsort,
 * arrops.py:332-335); here a set is sorted ONCE into a caller-owned store and then used by any number of
 * count/emit calls, or sorted on one rank and adopted by the others after a broadcast of its 8-byte keys and
 * 4-byte row ids (12 B/row instead of the 20 B/row of the raw columns, and no second sort).
 *
 *  1. bf_extent_measure per set and rank -> a small HOST array of bf_extent_words(n_groups) int64 words that
 *     merges by elementwise MAX (np.maximum between sets, all_reduce(MAX) across ranks).  Every set of one
 *     join must be prepared from the SAME merged words: they fix the key layout.
 *  2. bf_overlap_prepare_set (sorts; `presorted` = what bf_extent_measure reported for this very set), or
 *     bf_overlap_adopt_set + a copy of another rank's buffers into bf_sorted_set_buffers().
 *  3. bf_overlap_count_prepared -> caller allocates -> bf_overlap_emit (the same emit as above).
 *
 * own_range (host, 4 words or NULL) = (g_lo, start_lo, g_hi, start_hi): the call then only produces the part
 * of the result its shard owns -- A-block pairs and unpaired rows of the set-1 rows whose (group, start) lies
 * in [lo, hi) lexicographically, B-block pairs of the set-2 rows in that range.  Set-1 rows outside the range
 * are halo rows: they only appear as partners in the B block (a target's partners may start beyond the shard
 * boundary, arrops.py:341-342, 360-366).  With contiguous (group, start) ranges per shard, the reference's
 * row order of the whole join is, per group, the shards' A segments in shard order, then their B segments,
 * then their unpaired segments: bf_overlap_group_counts gives the segment sizes, bf_overlap_set_halo_rows the
 * global row numbers of the halo rows.  how=right/outer is refused with own_range (an unpaired target needs a
 * reduction over all shards). */
typedef struct bf_sorted_set {
  int64_t opaque[32];
} bf_sorted_set;
size_t bf_extent_words(int32_t n_groups);
size_t bf_extent_workspace_bytes(int32_t n_groups);
int bf_extent_measure(const int32_t* grp, const int64_t* start, const int64_t* end, int64_t n, int32_t n_groups,
                      void* ws, size_t ws_bytes, void* stream, int64_t* words_out, int32_t* presorted_out);
size_t bf_sorted_set_bytes(int64_t n, int32_t n_groups, int32_t adopt);
int bf_overlap_prepare_set(const int64_t* extent_words, const int32_t* grp, const int64_t* start, const int64_t* end,
                           int64_t n, int32_t n_groups, int32_t presorted, void* store, size_t store_bytes,
                           void* stream, bf_sorted_set* set_out);
int bf_overlap_adopt_set(const int64_t* extent_words, int64_t n, int32_t n_groups, void* store, size_t store_bytes,
                         void* stream, bf_sorted_set* set_out);
/* device pointers of the set's sorted keys (uint64[n]) and row ids (uint32[n]): what a broadcast moves */
int bf_sorted_set_buffers(const bf_sorted_set* set, void** keys_out, void** ids_out, int64_t* n_out);
/* optional, once per prepared set (after its buffers hold the sorted rows): the running maximum of end' in sorted
 * order (np.maximum.accumulate of the unpaired-row test) is computed into the set's own
 * store and reused by every bf_overlap_count_prepared that needs it, instead of once per call */
int bf_sorted_set_runmax(bf_sorted_set* set, void* stream);
size_t bf_overlap_prepared_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups, int32_t how);
/* grp1 / grp2: the group codes of the kept side(s) (how left: grp1), NULL otherwise or when n_groups == 1 */
int bf_overlap_count_prepared(const bf_sorted_set* set1, const bf_sorted_set* set2, const int32_t* grp1,
                              const int32_t* grp2, int32_t how, int32_t closed, const int64_t* own_range, void* ws,
                              size_t ws_bytes, bf_plan* plan, void* stream, int64_t* n_rows_out,
                              int64_t* n_pairs_out);
/* between _count and _emit: set-1 rows [n_own1, n1) are halo rows whose global row numbers are
 * halo_ids[row - n_own1] (device int64); rows below n_own1 get the row offset of bf_overlap_set_row_offsets */
int bf_overlap_set_halo_rows(bf_plan* plan, int64_t n_own1, const int64_t* halo_ids);
/* after a count call: counts_out (HOST, int64[4 * n_groups]) = per group the rows of its A block, B block,
 * unpaired set-1 rows, unpaired set-2 rows */
int bf_overlap_group_counts(const bf_plan* plan, void* ws, int64_t* counts_out, void* stream);

/* ---- overlap(how='left') with the rows in (row of set 1, row of set 2) order: what bioframe.overlap returns by
 * default for how='left' (keep_order=True, ops.py:455-456 and the sort_values of ops.py:549-550) when the frame
 * indexes order like the rows; -1 marks a row of set 1 without partner.  The pairs are produced in that order
 * directly -- the queries are visited in row order and never sorted, only the targets are -- instead of
 * sorting the pair list afterwards (bf_sort_pairs).  max_partners_out: the largest number of partners of one
 * row (lists above 32 are sorted by a single thread: a caller may prefer bf_overlap_* + bf_sort_pairs when it
 * is huge). */
size_t bf_overlap_ordered_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups);
int bf_overlap_ordered_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                             const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                             int32_t n_groups, void* ws, size_t ws_bytes, bf_plan* plan, void* stream,
                             int64_t* n_rows_out, int64_t* max_partners_out);
int bf_overlap_ordered_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out,
                            void* stream);

/* ---- keep_order: replaces `out_df.sort_values([index1, index2])` of ops.overlap (ops.py:549-550) for
 * frames whose indexes are monotonic (labels ordered like row positions).  Sorts the pairs in place by
 * (events1, events2); -1 sorts after every row number of its column (pandas: NA last).  max1 / max2 are
 * upper bounds of the row numbers in the two columns (they size the key). */
size_t bf_sort_pairs_workspace_bytes(int64_t n_rows);
int bf_sort_pairs(int64_t* events1, int64_t* events2, int64_t n_rows, int64_t max1, int64_t max2, void* ws,
                  size_t ws_bytes, void* stream);

/* ---- pair coordinates: replaces the gathers of return_overlap=True -- np.maximum(df1[sk1].values[events1],
 * df2[sk2].values[events2]) / np.minimum(...) in ops.overlap (ops.py:480-497) and ops.closest
 * (ops.py:1187-1194) -- and closest's return_distance arithmetic (ops.py:1207-1221), on the pair list
 * while it still is on the device.  For row i with a = events1[i], b = events2[i], both >= 0:
 *   overlap_start = max(start1[a], start2[b]), overlap_end = min(end1[a], end2[b]),
 *   distance = max(0, start1[a] - end2[b], start2[b] - end1[a]);
 * rows with a missing partner (-1) get 0 (the caller masks them as the reference does).  Any output may be
 * NULL. */
int bf_pair_coords(const int64_t* events1, const int64_t* events2, int64_t n_rows, const int64_t* start1,
                   const int64_t* end1, const int64_t* start2, const int64_t* end2, int64_t* overlap_start_out,
                   int64_t* overlap_end_out, int64_t* distance_out, void* stream);

/* ---- column gather: replaces `df.iloc[events]` of the frame assembly (ops.py:499-508, 1224-1234) for
 * fixed-width columns: out[i] = src[row_idx[i]] (itemsize 1, 2, 4 or 8 bytes, any type of that width:
 * numeric values, the integer codes of a categorical column, the mask bytes of a nullable column), 0 where
 * row_idx[i] is -1; valid_out (optional, uint8[n_rows]) = 1 where a row was taken -- value
 * buffer + validity, the Arrow layout of a nullable column. */
int bf_gather_rows(const void* src, int32_t itemsize, const int64_t* row_idx, int64_t n_rows, void* out,
                   uint8_t* valid_out, void* stream);

/* ---- group encoding: replaces the hash groupby that finds the (chromosome) groups of a frame --
 * df.groupby(by, observed=True).indices / .groups (ops.py:287-289, 631, 772, 983-985) -- for a categorical key
 * column, whose dictionary codes (int8 / int16 per row, -1 = null) go to the device as they are.
 * bf_group_codes_present: present_out (device int32[n_categories + 2]) gets 1 in slot 0 if a null occurs, in slot
 * c + 1 if category c occurs (the observed keys), in the last slot if any code lies outside [-1, n_categories).
 * bf_group_codes_apply: grp_out[i] = lut[codes[i] + 1] (lut: device int32[n_categories + 1], slot 0 = the rank of
 * null rows), i.e. the group rank every other entry point of this header takes. */
int bf_group_codes_present(const void* codes, int32_t itemsize, int64_t n, int32_t n_categories, int32_t* present_out,
                           void* stream);
int bf_group_codes_apply(const void* codes, int32_t itemsize, int64_t n, const int32_t* lut, int32_t n_categories,
                         int32_t* grp_out, void* stream);

/* ---- closest: replaces ops._closest_intidxs (ops.py:919-1040), i.e. the
 * per-group loop around arrops.closest_intervals (arrops.py:601-754) and
 * arrops._closest_intervals_nooverlap (arrops.py:506-598), plus the
 * np.setdiff1d of unmatched queries (ops.py:1027-1031).  Output rows, per group
 * in code order: for every query of the group in ascending row order its up to
 * k targets sorted by (distance, target row) -- overlaps have distance 0,
 * non-overlapping neighbours distance gap + 1 (arrops.py:724-730) -- then
 * (query, -1) for the group's queries without any candidate.
 *   flags: BF_CLOSEST_* bits; SELF = set 2 is set 1 (pass the same pointers):
 *          a query never reports itself as an overlap (arrops.py:652-657).
 *   along: device uint8[n1] or NULL; 0 marks a query on the "-" strand
 *          (direction_col, ops.py:1009-1012): upstream/downstream swap sides.
 * Ties among targets with equal end (left) / equal start (right) are resolved
 * in row order (stable); the reference uses an unstable argsort there
 * (arrops.py:562,580), so only tie-free inputs have a defined reference answer.
 * tie_breaking_col is not supported (raises in the reference, arrops.py:740). */
#define BF_CLOSEST_IGNORE_OVERLAPS 1
#define BF_CLOSEST_IGNORE_UPSTREAM 2
#define BF_CLOSEST_IGNORE_DOWNSTREAM 4
#define BF_CLOSEST_SELF 8
size_t bf_closest_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups);
int bf_closest_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                     const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                     int32_t n_groups, int32_t k, int32_t flags, const uint8_t* along, void* ws,
                     size_t ws_bytes, bf_plan* plan, void* stream, int64_t* n_rows_out,
                     int64_t* n_matched_out);
int bf_closest_emit(const bf_plan* plan, void* ws, int64_t* events1_out, int64_t* events2_out,
                    void* stream);

/* ---- cluster / merge: replaces the group loops of ops.cluster (ops.py:637-674)
 * and ops.merge (ops.py:776-810) around arrops.merge_intervals
 * (arrops.py:415-479).  Groups are visited in code order (= sorted key order,
 * `df.groupby(...).groups`), cluster ids run on across groups
 * (ops.py:654-656).  `has_min_dist` = 0 is min_dist=None (adjacent intervals are
 * NOT merged, arrops.py:470); otherwise borders are start > runmax + min_dist
 * (arrops.py:468).  `cluster_ids_out` (int64[n], input row order) may be NULL
 * (merge only needs the table).  bf_cluster_emit writes the cluster table
 * (group code, start, end, n_intervals -- np.bincount of ops.py:653; any may be
 * NULL) and, when `cluster_ids` + row outputs are given, cluster_start/end per
 * input row (ops.py:696-701). */
size_t bf_cluster_workspace_bytes(int64_t n, int32_t n_groups);
int bf_cluster_count(const int32_t* grp, const int64_t* start, const int64_t* end, int64_t n,
                     int32_t n_groups, int64_t min_dist, int32_t has_min_dist, void* ws,
                     size_t ws_bytes, bf_plan* plan, void* stream, int64_t* cluster_ids_out,
                     int64_t* n_clusters_out);
int bf_cluster_emit(const bf_plan* plan, void* ws, int32_t* cgrp_out, int64_t* cstart_out,
                    int64_t* cend_out, int64_t* ccount_out, const int64_t* cluster_ids,
                    int64_t* row_start_out, int64_t* row_end_out, void* stream);

/* ---- coverage: replaces ops.coverage (ops.py:842-916): merge(df2), left
 * overlap with clipped coordinates, per-query sum.  Computed without
 * enumerating pairs (prefix sums of merged lengths + two searches per query);
 * same integers out.  coverage_out: int64[n1] in input row order. */
size_t bf_coverage_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups);
int bf_coverage(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                int32_t n_groups, void* ws, size_t ws_bytes, void* stream, int64_t* coverage_out);

/* ---- count_overlaps: replaces ops.count_overlaps (ops.py:1371-1438: left overlap + groupby count) and
 * serves ops.setdiff (ops.py:1333-1368: rows with count 0).  counts_out: int64[n1] in input row order.
 * No pairs are enumerated (two binary searches per query over the targets sorted by start and by end'). */
size_t bf_count_overlaps_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups);
int bf_count_overlaps(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                      const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                      int32_t n_groups, int32_t closed, void* ws, size_t ws_bytes, void* stream,
                      int64_t* counts_out);

/* ---- complement: replaces the per-view-region loop of ops.complement
 * (ops.py:1644-1685) around arrops.complement_intervals (arrops.py:482-503).
 * `region` = code of the view region each (already clipped) interval belongs
 * to, regions in sorted name order; bounds_lo/hi = device int64[n_regions]
 * region start/end.  Output rows: per region in code order its gaps
 * (region code, start, end). */
size_t bf_complement_workspace_bytes(int64_t n, int32_t n_regions);
int bf_complement_count(const int32_t* region, const int64_t* start, const int64_t* end, int64_t n,
                        const int64_t* bounds_lo, const int64_t* bounds_hi, int32_t n_regions,
                        void* ws, size_t ws_bytes, bf_plan* plan, void* stream, int64_t* n_gaps_out);
int bf_complement_emit(const bf_plan* plan, void* ws, int32_t* region_out, int64_t* start_out,
                       int64_t* end_out, void* stream);

/* ---- subtract: replaces ops.subtract (ops.py:1243-1330), which the reference composes from
 * complement(df2) over [0, INT64_MAX) per chromosome (ops.py:1305-1309), a left overlap of df1 with those
 * gaps returning the clipped coordinates (ops.py:1310-1316, 480-497) and a filter of the rows without a
 * partner (ops.py:1318).  Fused here: merge of the targets + two probes per query, no pair list.
 * Output rows: for every row of set 1 in ascending row order (keep_order=True on a monotonic index) the
 * pieces of it not covered by set 2, left to right: row_out = row number in set 1, start/end = the piece,
 * sub_index_out = ordinal of the piece within its row ('sub_index' of return_index=True,
 * ops.py:1322-1328).  Rows that are covered entirely produce nothing.  row_out / sub_index_out may be NULL. */
size_t bf_subtract_workspace_bytes(int64_t n1, int64_t n2, int32_t n_groups);
int bf_subtract_count(const int32_t* grp1, const int64_t* start1, const int64_t* end1, int64_t n1,
                      const int32_t* grp2, const int64_t* start2, const int64_t* end2, int64_t n2,
                      int32_t n_groups, void* ws, size_t ws_bytes, bf_plan* plan, void* stream,
                      int64_t* n_rows_out);
int bf_subtract_emit(const bf_plan* plan, void* ws, int64_t* row_out, int64_t* start_out, int64_t* end_out,
                     int64_t* sub_index_out, void* stream);

/* ---- host download of row numbers: the last step of ops._overlap_intidxs / ops._closest_intidxs as the
 * reference's callers see it -- int64 numpy arrays of row positions on the HOST (ops.py:338, 1040).  A call
 * through host buffers is bound by PCIe, and row numbers of sets below 2^31 rows carry 4 bytes: they cross the
 * bus as int32 and are widened into `host_dst` by `n_threads` host threads (streaming stores) while the next
 * chunk is in flight.  dev_src: device int64[n] with every value in [-1, 2^31 - 1), 16-byte aligned (else
 * BF_ERR_RANGE / BF_ERR_ARG); host_dst: ordinary (pageable) host int64[n]; dev_stage: device scratch of
 * bf_download_stage_bytes(chunk_rows) bytes; pinned_stage: page-locked host scratch of at least
 * 8 * chunk_rows bytes.  Synchronises `stream`; returns when host_dst is complete.
 *   When host_dst itself is page-locked and the call spans at least 8 chunks, the call times the widening and the
 * copy of its first chunks; on a host whose threads are the slower side (the boxes differ by 5x) the tail of the
 * array then travels as plain int64 copies straight into host_dst on a second stream, a few slices per narrow
 * chunk, until the two lanes meet -- same result, 8 instead of 4 bytes per row over the bus for that part, no host
 * work.  Environment BF_DOWNLOAD_DIRECT=0 disables the second lane, =1 forces it (tests).
 * bf_download_direct_rows(): rows of the calling thread's last bf_download_rows call that took it. */
size_t bf_download_stage_bytes(int64_t chunk_rows);
int64_t bf_download_direct_rows(void);
int bf_download_rows(const int64_t* dev_src, int64_t n, int64_t* host_dst, void* dev_stage,
                     size_t dev_stage_bytes, void* pinned_stage, size_t pinned_bytes, int32_t n_threads,
                     void* stream);

/* ---- BED ingest: replaces `read_table(path, schema="bed3" ...)` (io/fileops.py:42-84: pandas.read_csv with
 * sep="\t", header=None) plus the host-side encoding of the chromosome column (ops.py:287-289) for the three
 * columns the path consumes: tab-separated text on the device -> chromosome code, start, end.  Every
 * non-blank line is a record `<name> TAB <int> TAB <int> [TAB anything]`, lines end in "\n" or "\r\n" (the last
 * one may be unterminated); anything else is an error naming the line.  text: device uint8[n_bytes], 16-byte
 * aligned.
 *   bf_bed_scan   -> number of lines.
 *   bf_bed_parse  -> number of records (n_rows_out) and the distinct names: for name k (k < *n_names_out <=
 *                    32768) the byte offset of an occurrence (HOST name_first_out[k]), its length (HOST
 *                    name_len_out[k]) and its table slot (HOST name_slot_out[k]); the three arrays must hold
 *                    32768 entries.
 *   bf_bed_emit   -> rows in file order; slot_code: HOST int32[65536], the code the caller assigns to every
 *                    slot reported by bf_bed_parse (e.g. the rank of the name in sorted order).  Outputs may
 *                    be NULL. */
size_t bf_bed_scan_workspace_bytes(int64_t n_bytes);
int bf_bed_scan(const uint8_t* text, int64_t n_bytes, void* ws, size_t ws_bytes, void* stream, int64_t* n_lines_out);
size_t bf_bed_parse_workspace_bytes(int64_t n_bytes, int64_t n_lines);
int bf_bed_parse(const uint8_t* text, int64_t n_bytes, int64_t n_lines, void* ws, size_t ws_bytes, bf_plan* plan,
                 void* stream, int64_t* n_rows_out, int32_t* n_names_out, int64_t* name_first_out,
                 int32_t* name_len_out, int32_t* name_slot_out);
int bf_bed_emit(const bf_plan* plan, void* ws, const int32_t* slot_code, int32_t* chrom_code_out,
                int64_t* start_out, int64_t* end_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIOFRAME_B200_H */
