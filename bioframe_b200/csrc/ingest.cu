// bf_bed_*: tab-separated interval text (BED) -> device columns (chromosome code, start, end), the three
// columns every operation of the path consumes.  Replaces `read_table(path, schema="bed3"...)`
// (src/bioframe/io/fileops.py:42-84 = pandas.read_csv(sep="\t", header=None)) followed by the host-side
// group encoding of the chromosome column (ops.py:287-289 / SURVEY.md 8a a14) for these columns: the text goes to
// the device once and the chromosome names come back dictionary-encoded.
//
//   bf_bed_scan    newline census: per 32 KB tile the number of line ends  -> spine scan -> line count
//   bf_bed_parse   line starts (same census, now writing), then one thread per line: blank lines are skipped
//                  (pandas skip_blank_lines), field 0 is hashed (FNV-1a 64) into a small open-addressing table
//                  of distinct names, fields 1 and 2 are parsed as decimal int64; kept lines get their output
//                  row through an in-tile scan + spine
//   host           reads the few distinct names (offset, length of a first occurrence), orders them, sends
//                  back slot -> code
//   bf_bed_emit    rows in file order: code, start, end
// Byte work, HBM-bound by design (the text is read three times: 3 B per byte); no tensor cores.
#include <vector>

#include "common.cuh"
#include "scan.cuh"
#include "../../include/bioframe_b200.h"

namespace bf {

constexpr int BED_CHUNK = 16;                              // bytes per thread and step (one 16-byte load)
constexpr int BED_TILE = SC_THREADS * BED_CHUNK * 8;       // 32 KB of text per block
constexpr u32 BED_SLOTS = 1u << 16;                        // distinct chromosome names (open addressing)
constexpr u32 BED_MAX_NAMES = BED_SLOTS / 2;

static inline i64 bed_tiles(i64 n_bytes) { return ceil_div(n_bytes, BED_TILE); }

struct BedSlot {
  u64 hash;    // 0 = empty (hashes are forced non-zero)
  u64 first;   // smallest byte offset of a line carrying this name
  u32 len;     // length of the name (set by every writer: equal for equal names)
  u32 check;   // second, independent hash: a mismatch means two names share `hash`
};

struct BedWS {
  u64* spine;       // [tiles + 2] line ends per text tile
  i64* line_start;  // [n_lines + 1]
  i64 *t_start, *t_end;  // [n_lines] parsed fields, by line
  u32* t_slot;      // [n_lines] name slot, 0xffffffff = blank line
  u32* loc;         // [n_lines] output row inside the line tile
  u64* row_spine;   // [line tiles + 2]
  BedSlot* table;   // [BED_SLOTS]
  i32* remap;       // [BED_SLOTS] slot -> code (filled by the host between parse and emit)
  i64* hdr;         // [0] lines, [1] rows, [2] first bad line + 1 (0 = none), [3] distinct names
};
static void layout_bed(Arena& a, i64 n_bytes, i64 n_lines, BedWS* w) {
  w->spine = a.take<u64>(bed_tiles(n_bytes) + 2);
  w->hdr = a.take<i64>(8);
  w->table = a.take<BedSlot>(BED_SLOTS);
  w->remap = a.take<i32>(BED_SLOTS);
  w->line_start = a.take<i64>(n_lines + 2);
  w->t_start = a.take<i64>(n_lines + 1);
  w->t_end = a.take<i64>(n_lines + 1);
  w->t_slot = a.take<u32>(n_lines + 1);
  w->loc = a.take<u32>(n_lines + 1);
  w->row_spine = a.take<u64>(scan_tiles((u64)(n_lines > 0 ? n_lines : 1)) + 2);
}

struct BedPlan {
  u64 magic;
  i64 n_bytes, n_lines, n_rows;
  i32 n_names, ready;
  const u8* text;
};
static_assert(sizeof(BedPlan) <= sizeof(bf_plan), "bf_plan too small");
constexpr u64 BED_MAGIC = 0x6266426564496e67ull;

__device__ __forceinline__ u32 newline_mask16(const u8* __restrict__ text, i64 at, i64 n) {
  u32 m = 0;
  if (at + BED_CHUNK <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(text + at);
    const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (((w[k] >> (8 * b)) & 0xffu) == '\n') m |= 1u << (4 * k + b);
  } else {
    for (int b = 0; b < BED_CHUNK && at + b < n; ++b)
      if (text[at + b] == '\n') m |= 1u << b;
  }
  return m;
}

// pass 1: line ends per tile (a last line without '\n' counts too)
__global__ void __launch_bounds__(SC_THREADS) bed_census_kernel(const u8* __restrict__ text, i64 n, u64* __restrict__ spine) {
  __shared__ u64 s_tmp[8];
  const i64 base = (i64)blockIdx.x * BED_TILE;
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const i64 at = base + ((i64)j * SC_THREADS + threadIdx.x) * BED_CHUNK;
    if (at < n) c += __popc(newline_mask16(text, at, n));
  }
  c = block_sum_u64(c, s_tmp);
  if (threadIdx.x == 0) {
    if (base <= n - 1 && n - 1 < base + BED_TILE && text[n - 1] != '\n') ++c;  // unterminated last line
    spine[blockIdx.x] = c;
  }
}

// pass 2: line k + 1 starts after the k-th line end; line 0 starts at byte 0
__global__ void __launch_bounds__(SC_THREADS)
    bed_line_starts_kernel(const u8* __restrict__ text, i64 n, const u64* __restrict__ spine, i64* __restrict__ line_start) {
  __shared__ u64 s_tmp[8];
  const i64 base = (i64)blockIdx.x * BED_TILE;
  // thread t owns 8 consecutive 16-byte chunks = 128 consecutive bytes
  const i64 mine = base + (i64)threadIdx.x * (BED_CHUNK * 8);
  u32 m[8];
  u64 c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const i64 at = mine + j * BED_CHUNK;
    m[j] = at < n ? newline_mask16(text, at, n) : 0u;
    c += __popc(m[j]);
  }
  u64 tot;
  u64 k = spine[blockIdx.x] + block_excl_sum_u64(c, s_tmp, &tot);  // line ends before this thread's bytes
  if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u32 bits = m[j];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      line_start[++k] = mine + j * BED_CHUNK + b + 1;
    }
  }
}

__device__ __forceinline__ bool bed_parse_int(const u8* __restrict__ text, i64& p, i64 end, i64* out) {
  bool neg = false;
  if (p < end && (text[p] == '-' || text[p] == '+')) neg = text[p++] == '-';
  if (p >= end || text[p] < '0' || text[p] > '9') return false;
  u64 v = 0;
  int digits = 0;
  while (p < end && text[p] >= '0' && text[p] <= '9') {
    v = v * 10 + (u64)(text[p] - '0');
    ++p;
    if (++digits > 19) return false;
  }
  if (v > (neg ? 0x8000000000000000ull : 0x7fffffffffffffffull)) return false;
  *out = neg ? (i64)(0 - v) : (i64)v;
  return true;
}

// pass 3: one thread per line
__global__ void __launch_bounds__(SC_THREADS)
    bed_parse_kernel(const u8* __restrict__ text, i64 n_bytes, const i64* __restrict__ line_start, i64 n_lines,
                     BedSlot* __restrict__ table, i64* __restrict__ t_start, i64* __restrict__ t_end,
                     u32* __restrict__ t_slot, u32* __restrict__ loc, u64* __restrict__ row_spine, i64* __restrict__ hdr) {
  __shared__ u64 s_tmp[8];
  const i64 base = (i64)blockIdx.x * SC_TILE + (i64)threadIdx.x * SC_IPT;
  u32 keep = 0;  // bit j: line base + j yields a row
  u64 mine = 0;
#pragma unroll 1
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 ln = base + j;
    if (ln >= n_lines) break;
    i64 p = line_start[ln];
    i64 end = ln + 1 < n_lines ? line_start[ln + 1] - 1 : n_bytes;  // the '\n' itself is not part of the line
    if (ln + 1 >= n_lines && end > p && text[end - 1] == '\n') --end;
    if (end > p && text[end - 1] == '\r') --end;
    u32 slot = 0xffffffffu;
    if (end > p) {  // not blank
      const i64 name0 = p;
      u64 h = 1469598103934665603ull;
      u32 h2 = 2166136261u;
      while (p < end && text[p] != '\t') {
        h = (h ^ text[p]) * 1099511628211ull;
        h2 = (h2 ^ text[p]) * 16777619u + 0x9e3779b9u;
        ++p;
      }
      const u32 len = (u32)(p - name0);
      i64 a = 0, b = 0;
      bool ok = len > 0 && p < end;
      if (ok) {
        ++p;  // tab
        ok = bed_parse_int(text, p, end, &a) && p < end && text[p] == '\t';
      }
      if (ok) {
        ++p;
        ok = bed_parse_int(text, p, end, &b) && (p == end || text[p] == '\t');
      }
      if (!ok) {
        atomicMin(reinterpret_cast<unsigned long long*>(hdr + 2), (unsigned long long)(ln + 1));
      } else {
        if (h == 0) h = 1;
        if (h2 == 0) h2 = 1;
        u32 s = (u32)(h >> 17) & (BED_SLOTS - 1);
        for (u32 probe = 0; probe < BED_SLOTS; ++probe, s = (s + 1) & (BED_SLOTS - 1)) {
          u64 seen = table[s].hash;
          if (seen == 0) seen = atomicCAS(reinterpret_cast<unsigned long long*>(&table[s].hash), 0ull, h);
          if (seen == 0 || seen == h) {
            // a genome has a few dozen names and every line lands on one of them: once a slot is set up,
            // plain loads see that nothing is left to do (no atomics on the hot addresses)
            u32 other = *reinterpret_cast<volatile u32*>(&table[s].check);
            if (other == 0) other = atomicCAS(&table[s].check, 0u, h2);
            if (other != 0 && other != h2) continue;  // same 64-bit hash, another name: that one owns the slot, probe on
            if (*reinterpret_cast<volatile u64*>(&table[s].first) > (u64)name0)
              atomicMin(reinterpret_cast<unsigned long long*>(&table[s].first), (unsigned long long)name0);
            if (*reinterpret_cast<volatile u32*>(&table[s].len) != len) table[s].len = len;
            slot = s;
            break;
          }
        }
        if (slot == 0xffffffffu) atomicMin(reinterpret_cast<unsigned long long*>(hdr + 2), (unsigned long long)(ln + 1));
        t_start[ln] = a;
        t_end[ln] = b;
        if (slot != 0xffffffffu) {
          keep |= 1u << j;
          ++mine;
        }
      }
    }
    t_slot[ln] = slot;
  }
  u64 tot;
  u64 run = block_excl_sum_u64(mine, s_tmp, &tot);
  if (threadIdx.x == 0) row_spine[blockIdx.x] = tot;
#pragma unroll
  for (int j = 0; j < SC_IPT; ++j) {
    const i64 ln = base + j;
    if (ln < n_lines) loc[ln] = (u32)run;
    run += (keep >> j) & 1u;
  }
}

// pass 4: the names, byte for byte.  A slot is identified by a 64-bit + 32-bit fingerprint of the name; two
// different names with the same 96 bits would silently become one chromosome.  Every line compares its name with
// the first occurrence recorded in its slot (a handful of bytes per line, the same cache lines the parse
// just read); a mismatch is reported as what it is, a fingerprint collision, with the line.
__global__ void __launch_bounds__(256)
    bed_verify_names_kernel(const u8* __restrict__ text, const i64* __restrict__ line_start, i64 n_lines,
                            const BedSlot* __restrict__ table, const u32* __restrict__ t_slot, i64* __restrict__ hdr) {
  const i64 ln = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= n_lines) return;
  const u32 slot = t_slot[ln];
  if (slot == 0xffffffffu) return;
  const u64 first = table[slot].first;
  const u32 len = table[slot].len;
  const i64 p = line_start[ln];
  if ((u64)p == first) return;
  bool same = true;
  for (u32 i = 0; i < len && same; ++i) same = text[p + i] == text[first + i];
  same = same && text[p + len] == '\t';
  if (!same) atomicMin(reinterpret_cast<unsigned long long*>(hdr + 3), (unsigned long long)(ln + 1));
}

__global__ void bed_table_init_kernel(BedSlot* table) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < BED_SLOTS) {
    table[i].hash = 0;
    table[i].first = ~0ull;
    table[i].len = 0;
    table[i].check = 0;
  }
}
__global__ void __launch_bounds__(256)
    bed_emit_kernel(const i64* __restrict__ t_start, const i64* __restrict__ t_end, const u32* __restrict__ t_slot,
                    const u32* __restrict__ loc, const u64* __restrict__ row_spine, const i32* __restrict__ remap,
                    i64 n_lines, i32* __restrict__ code_out, i64* __restrict__ start_out, i64* __restrict__ end_out) {
  const i64 ln = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= n_lines) return;
  const u32 slot = t_slot[ln];
  if (slot == 0xffffffffu) return;
  const i64 row = (i64)row_spine[ln / SC_TILE] + loc[ln];
  if (code_out) code_out[row] = remap[slot];
  if (start_out) start_out[row] = t_start[ln];
  if (end_out) end_out[row] = t_end[ln];
}

}  // namespace bf

extern "C" {

size_t bf_bed_scan_workspace_bytes(int64_t n_bytes) {
  if (n_bytes < 0) return 0;
  return bf::align_up((size_t)(bf::bed_tiles(n_bytes) + 2) * sizeof(u64)) + 256;
}
int bf_bed_scan(const uint8_t* text, int64_t n_bytes, void* ws, size_t ws_bytes, void* stream, int64_t* n_lines_out) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_bytes < 0 || !n_lines_out) {
    set_error("bf_bed_scan: bad argument");
    return BF_ERR_ARG;
  }
  *n_lines_out = 0;
  if (n_bytes == 0) return BF_OK;
  if (!text || !ws || ws_bytes < bf_bed_scan_workspace_bytes(n_bytes) || (reinterpret_cast<uintptr_t>(text) & 15)) {
    set_error("bf_bed_scan: null / unaligned text or workspace too small");
    return BF_ERR_WORKSPACE;
  }
  u64* spine = reinterpret_cast<u64*>(ws);
  const i64 tiles = bed_tiles(n_bytes);
  BF_LAUNCH(PC_GATHER, n_bytes, st, bed_census_kernel<<<(unsigned)tiles, SC_THREADS, 0, st>>>(text, n_bytes, spine));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(spine, (u64)tiles, spine + tiles + 1));
  u64 total = 0;
  BF_CUDA(cudaMemcpyAsync(&total, spine + tiles + 1, sizeof(u64), cudaMemcpyDeviceToHost, st));
  BF_CUDA(cudaStreamSynchronize(st));
  *n_lines_out = (int64_t)total;
  return BF_OK;
}

size_t bf_bed_parse_workspace_bytes(int64_t n_bytes, int64_t n_lines) {
  if (n_bytes < 0 || n_lines < 0) return 0;
  bf::Arena a(nullptr);
  bf::BedWS w;
  bf::layout_bed(a, n_bytes, n_lines, &w);
  return a.off;
}
int bf_bed_parse(const uint8_t* text, int64_t n_bytes, int64_t n_lines, void* ws, size_t ws_bytes, bf_plan* plan_,
                 void* stream, int64_t* n_rows_out, int32_t* n_names_out, int64_t* name_first_out,
                 int32_t* name_len_out, int32_t* name_slot_out) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  BedPlan* plan = (BedPlan*)plan_;
  if (!plan || n_bytes < 0 || n_lines < 0 || !n_rows_out || !n_names_out) {
    set_error("bf_bed_parse: bad argument");
    return BF_ERR_ARG;
  }
  plan->magic = 0;
  plan->ready = 0;
  plan->n_bytes = n_bytes;
  plan->n_lines = n_lines;
  plan->n_rows = 0;
  plan->n_names = 0;
  plan->text = text;
  *n_rows_out = 0;
  *n_names_out = 0;
  if (n_lines == 0) {
    plan->magic = BED_MAGIC;
    plan->ready = 1;
    return BF_OK;
  }
  if (n_lines >= (1ll << 31)) {
    set_error("bf_bed_parse: more than 2^31-1 lines");
    return BF_ERR_RANGE;
  }
  Arena measure(nullptr);
  BedWS w;
  layout_bed(measure, n_bytes, n_lines, &w);
  if (!text || !ws || ws_bytes < measure.off || !name_first_out || !name_len_out || !name_slot_out) {
    set_error("bf_bed_parse: null pointer or workspace %zu bytes < required %zu", ws_bytes, measure.off);
    return BF_ERR_WORKSPACE;
  }
  Arena arena(ws);
  layout_bed(arena, n_bytes, n_lines, &w);
  const i64 tiles = bed_tiles(n_bytes);
  const unsigned ltiles = (unsigned)scan_tiles((u64)n_lines);
  BF_CUDA(cudaMemsetAsync(w.hdr, 0, 8 * sizeof(i64), st));
  BF_CUDA(cudaMemsetAsync(w.hdr + 2, 0xff, 2 * sizeof(i64), st));  // first bad line, first colliding line: none
  BF_LAUNCH(PC_GATHER, 0, st, bed_table_init_kernel<<<BED_SLOTS / 256, 256, 0, st>>>(w.table));
  // the census again (the workspace of bf_bed_scan is not kept), now followed by the line starts
  BF_LAUNCH(PC_GATHER, n_bytes, st, bed_census_kernel<<<(unsigned)tiles, SC_THREADS, 0, st>>>(text, n_bytes, w.spine));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.spine, (u64)tiles, reinterpret_cast<u64*>(w.hdr)));
  BF_LAUNCH(PC_GATHER, n_bytes + 8 * n_lines, st,
            bed_line_starts_kernel<<<(unsigned)tiles, SC_THREADS, 0, st>>>(text, n_bytes, w.spine, w.line_start));
  BF_LAUNCH(PC_GATHER, n_bytes + 32 * n_lines, st,
            bed_parse_kernel<<<ltiles, SC_THREADS, 0, st>>>(text, n_bytes, w.line_start, n_lines, w.table, w.t_start, w.t_end,
                                                            w.t_slot, w.loc, w.row_spine, w.hdr));
  BF_LAUNCH(PC_SCAN, 0, st, spine_scan_kernel<false><<<1, 1024, 0, st>>>(w.row_spine, (u64)ltiles, reinterpret_cast<u64*>(w.hdr + 1)));
  BF_LAUNCH(PC_GATHER, 16 * n_lines, st,
            bed_verify_names_kernel<<<(unsigned)ceil_div(n_lines, 256), 256, 0, st>>>(text, w.line_start, n_lines, w.table,
                                                                                     w.t_slot, w.hdr));
  BF_CUDA(cudaStreamSynchronize(st));
  i64 hdr[4];
  BF_CUDA(cudaMemcpy(hdr, w.hdr, sizeof(hdr), cudaMemcpyDeviceToHost));
  if ((i64)hdr[0] != n_lines) {
    set_error("bf_bed_parse: n_lines = %lld, the text has %lld lines (pass the count bf_bed_scan returned)",
              (long long)n_lines, (long long)hdr[0]);
    return BF_ERR_ARG;
  }
  if (hdr[2] != -1) {
    set_error("bf_bed_parse: line %lld is not <name> TAB <integer> TAB <integer> [TAB ...] (or more than %u distinct names)",
              (long long)hdr[2], BED_MAX_NAMES);
    return BF_ERR_ARG;
  }
  if (hdr[3] != -1) {
    set_error("bf_bed_parse: the chromosome name of line %lld has the same 96-bit fingerprint as a different name "
              "(hash collision; the names are compared byte for byte, nothing was merged)", (long long)hdr[3]);
    return BF_ERR_RANGE;
  }
  // host copy of the table: BED_SLOTS x 24 bytes = 1.5 MB
  std::vector<BedSlot> h_table(BED_SLOTS);
  BF_CUDA(cudaMemcpy(h_table.data(), w.table, sizeof(BedSlot) * BED_SLOTS, cudaMemcpyDeviceToHost));
  i32 k = 0;
  for (u32 s = 0; s < BED_SLOTS; ++s)
    if (h_table[s].hash != 0) {
      if (k >= (i32)BED_MAX_NAMES) {
        set_error("bf_bed_parse: more than %u distinct names", BED_MAX_NAMES);
        return BF_ERR_RANGE;
      }
      name_first_out[k] = (int64_t)h_table[s].first;
      name_len_out[k] = (int32_t)h_table[s].len;
      name_slot_out[k] = (int32_t)s;
      ++k;
    }
  plan->n_rows = hdr[1];
  plan->n_names = k;
  plan->magic = BED_MAGIC;
  plan->ready = 1;
  *n_rows_out = hdr[1];
  *n_names_out = k;
  return BF_OK;
}

int bf_bed_emit(const bf_plan* plan_, void* ws, const int32_t* slot_code, int32_t* chrom_code_out, int64_t* start_out,
                int64_t* end_out, void* stream) {
  using namespace bf;
  cudaStream_t st = (cudaStream_t)stream;
  const BedPlan* plan = (const BedPlan*)plan_;
  if (!plan || plan->magic != BED_MAGIC || !plan->ready) {
    set_error("bf_bed_emit: plan was not produced by a successful bf_bed_parse");
    return BF_ERR_STATE;
  }
  if (plan->n_rows == 0) return BF_OK;
  if (!ws || (chrom_code_out && !slot_code)) {
    set_error("bf_bed_emit: null pointer");
    return BF_ERR_ARG;
  }
  Arena arena(ws);
  BedWS w;
  layout_bed(arena, plan->n_bytes, plan->n_lines, &w);
  // slot_code: HOST int32[65536], code of every occupied slot (the caller orders the names)
  if (chrom_code_out) BF_CUDA(cudaMemcpyAsync(w.remap, slot_code, sizeof(i32) * BED_SLOTS, cudaMemcpyHostToDevice, st));
  BF_LAUNCH(PC_GATHER, 28 * plan->n_lines + 20 * plan->n_rows, st,
            bed_emit_kernel<<<(unsigned)ceil_div(plan->n_lines, 256), 256, 0, st>>>(w.t_start, w.t_end, w.t_slot, w.loc,
                                                                                   w.row_spine, w.remap, plan->n_lines,
                                                                                   chrom_code_out, (i64*)start_out,
                                                                                   (i64*)end_out));
  BF_CUDA(cudaStreamSynchronize(st));  // slot_code may be freed by the caller
  return BF_OK;
}
}
