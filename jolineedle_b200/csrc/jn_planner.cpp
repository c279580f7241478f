// Host-side planner of supervised episodes: the host half of NeedleSimpleEnv.generate_sample
// (reference: src/env/simple_env.py:378-441,481-629,666-718) in C++.
//
// The reference makes its random decisions with a numpy Generator (PCG64 seeded through
// SeedSequence), python's global `random` (MT19937) and -- implicitly -- the iteration order of
// CPython `set` objects holding (y, x) tuples.  To reproduce a seeded reference trajectory bit
// for bit without paying ~300 us of interpreter time per episode, this file restates those three
// mechanisms from their published algorithms:
//
//   * numpy.random.SeedSequence -> PCG64 (XSL-RR 128/64) -> next_uint32 buffering ->
//     Generator.integers (Lemire, 32-bit path), Generator.choice(int) (== integers(0, n)),
//     Generator.binomial (inversion for n*p <= 30, BTPE above)        [numpy >= 1.17 stream]
//   * random.choice -> _randbelow_with_getrandbits -> MT19937 genrand_uint32  [CPython 3.x]
//   * set: open addressing, LINEAR_PROBES = 9, perturb shift 5, growth at fill*5 >= mask*3,
//     dummy entries on removal, set_merge fast paths; tuple hash = the xxHash-style combiner of
//     CPython >= 3.8                                                   [CPython 3.8 - 3.13]
//
// tests/test_planner_cpu.py checks the result against the pure-python planner (which calls the
// real numpy / random / set) on thousands of random episodes; the python planner stays the
// fallback for inputs this file does not cover (float boxes, seeds above 2^64).
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "../../include/jolineedle_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// numpy: SeedSequence + PCG64 + Generator methods
// ------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;

struct NumpyRng {
  u128 state = 0, inc = 0;
  bool has32 = false;
  uint32_t buf32 = 0;

  static uint32_t hashmix(uint32_t v, uint32_t& hc) {
    v ^= hc;
    hc *= 0x931e8875u;
    v *= hc;
    v ^= v >> 16;
    return v;
  }
  static uint32_t mix(uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
    r ^= r >> 16;
    return r;
  }

  void seed(const uint32_t* entropy, int n_words) {
    uint32_t pool[4], hc = 0x43b0d7e5u;
    for (int i = 0; i < 4; ++i) pool[i] = hashmix(i < n_words ? entropy[i] : 0u, hc);
    for (int s = 0; s < 4; ++s)
      for (int d = 0; d < 4; ++d)
        if (s != d) pool[d] = mix(pool[d], hashmix(pool[s], hc));
    for (int s = 4; s < n_words; ++s)
      for (int d = 0; d < 4; ++d) pool[d] = mix(pool[d], hashmix(entropy[s], hc));
    // generate_state(4, uint64) = 8 uint32 words, low word first
    uint32_t w[8], hb = 0x8b51f9ddu;
    for (int i = 0; i < 8; ++i) {
      uint32_t v = pool[i & 3];
      v ^= hb;
      hb *= 0x58f38dedu;
      v *= hb;
      v ^= v >> 16;
      w[i] = v;
    }
    uint64_t v64[4];
    for (int i = 0; i < 4; ++i) v64[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    const u128 initstate = ((u128)v64[0] << 64) | v64[1], initseq = ((u128)v64[2] << 64) | v64[3];
    state = 0;
    inc = (initseq << 1) | 1;
    step();
    state += initstate;
    step();
    has32 = false;
  }
  void seed_u64(uint64_t s) {
    uint32_t e[2] = {(uint32_t)s, (uint32_t)(s >> 32)};
    seed(e, e[1] ? 2 : 1);
  }
  void step() {
    const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
    state = state * mult + inc;
  }
  uint64_t next64() {
    step();
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned r = (unsigned)(hi >> 58);
    return (x >> r) | (x << ((64 - r) & 63));
  }
  uint32_t next32() {
    if (has32) {
      has32 = false;
      return buf32;
    }
    const uint64_t n = next64();
    has32 = true;
    buf32 = (uint32_t)(n >> 32);
    return (uint32_t)n;
  }
  double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
  // Generator.integers(low, high) for spans below 2^32 (also Generator.choice(int))
  int64_t integers(int64_t low, int64_t high) {
    const uint32_t rng = (uint32_t)(high - low - 1);
    if (rng == 0) return low;
    const uint32_t excl = rng + 1;
    uint64_t m = (uint64_t)next32() * excl;
    uint32_t left = (uint32_t)m;
    if (left < excl) {
      const uint32_t thr = (0xFFFFFFFFu - rng) % excl;
      while (left < thr) {
        m = (uint64_t)next32() * excl;
        left = (uint32_t)m;
      }
    }
    return low + (int64_t)(m >> 32);
  }
  // Generator.binomial(n, p) for p <= 0.5 (numpy's random_binomial: inversion up to n * p = 30, BTPE above)
  int64_t binomial(int64_t n, double p) {
    if (n == 0 || p == 0.0) return 0;
    if (p * (double)n > 30.0) return binomial_btpe(n, p);
    const double q = 1.0 - p, qn = std::exp((double)n * std::log(q)), np = (double)n * p;
    const double bound = std::fmin((double)n, np + 10.0 * std::sqrt(np * q + 1));
    int64_t X = 0;
    double px = qn, U = next_double();
    while (U > px) {
      X++;
      if ((double)X > bound) {
        X = 0;
        px = qn;
        U = next_double();
      } else {
        U -= px;
        px = ((double)(n - X + 1) * p * px) / ((double)X * q);
      }
    }
    return X;
  }
  // BTPE (Kachitvichyanukul & Schmeiser 1988) exactly as numpy's random_binomial_btpe draws it: two doubles
  // per trial (u scaled by p4, then v), the same four regions, the same squeeze / final acceptance tests in
  // the same floating-point order.  p <= 0.5.
  int64_t binomial_btpe(int64_t n, double p) {
    const double r = std::fmin(p, 1.0 - p), q = 1.0 - r;
    const double fm = (double)n * r + r;
    const int64_t m = (int64_t)std::floor(fm);
    const double p1 = std::floor(2.195 * std::sqrt((double)n * r * q) - 4.6 * q) + 0.5;
    const double xm = (double)m + 0.5, xl = xm - p1, xr = xm + p1;
    const double c = 0.134 + 20.5 / (15.3 + (double)m);
    double a = (fm - xl) / (fm - xl * r);
    const double laml = a * (1.0 + a / 2.0);
    a = (xr - fm) / (xr * q);
    const double lamr = a * (1.0 + a / 2.0);
    const double p2 = p1 * (1.0 + 2.0 * c), p3 = p2 + c / laml, p4 = p3 + c / lamr;
    const double nrq = (double)n * r * q;
    for (;;) {
      double u = next_double() * p4, v = next_double();
      int64_t y;
      if (u <= p1) {  // triangular centre: accepted at once
        y = (int64_t)std::floor(xm - p1 * v + u);
        return y;
      }
      if (u <= p2) {  // parallelogram
        const double x = xl + (u - p1) / c;
        v = v * c + 1.0 - std::fabs((double)m - x + 0.5) / p1;
        if (v > 1.0) continue;
        y = (int64_t)std::floor(x);
      } else if (u <= p3) {  // left exponential tail
        y = (int64_t)std::floor(xl + std::log(v) / laml);
        if (y < 0 || v == 0.0) continue;
        v = v * (u - p2) * laml;
      } else {  // right exponential tail
        y = (int64_t)std::floor(xr - std::log(v) / lamr);
        if (y > n || v == 0.0) continue;
        v = v * (u - p3) * lamr;
      }
      const int64_t k = y > m ? y - m : m - y;
      if (!(k > 20 && (double)k < nrq / 2.0 - 1)) {  // explicit evaluation of f(y) / f(m)
        const double s = r / q, aa = s * (double)(n + 1);
        double F = 1.0;
        if (m < y) {
          for (int64_t i = m + 1; i <= y; i++) F *= (aa / (double)i - s);
        } else if (m > y) {
          for (int64_t i = y + 1; i <= m; i++) F /= (aa / (double)i - s);
        }
        if (v > F) continue;
        return y;
      }
      // squeezes on log(v), then the Stirling-corrected bound
      const double kk = (double)k;
      const double rho = (kk / nrq) * ((kk * (kk / 3.0 + 0.625) + 0.16666666666666666) / nrq + 0.5);
      const double t = -kk * kk / (2 * nrq);
      const double A = std::log(v);
      if (A < t - rho) return y;
      if (A > t + rho) continue;
      const double x1 = (double)(y + 1), f1 = (double)(m + 1), z = (double)(n + 1 - m), w = (double)(n - y + 1);
      const double x2 = x1 * x1, f2 = f1 * f1, z2 = z * z, w2 = w * w;
      if (A > (xm * std::log(f1 / x1) + ((double)(n - m) + 0.5) * std::log(z / w) +
               (double)(y - m) * std::log(w * r / (x1 * q)) +
               (13680. - (462. - (132. - (99. - 140. / f2) / f2) / f2) / f2) / f1 / 166320. +
               (13680. - (462. - (132. - (99. - 140. / z2) / z2) / z2) / z2) / z / 166320. +
               (13680. - (462. - (132. - (99. - 140. / x2) / x2) / x2) / x2) / x1 / 166320. +
               (13680. - (462. - (132. - (99. - 140. / w2) / w2) / w2) / w2) / w / 166320.))
        continue;
      return y;
    }
  }
};

// ------------------------------------------------------------------------------------------
// CPython: MT19937 behind random.choice
// ------------------------------------------------------------------------------------------
struct PyRandom {
  uint32_t* mt;  // 624 words
  uint32_t* idx; // position
  uint32_t genrand() {
    constexpr int N = 624, M = 397;
    if (*idx >= (uint32_t)N) {
      int kk;
      uint32_t y;
      for (kk = 0; kk < N - M; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      for (; kk < N - 1; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      *idx = 0;
    }
    uint32_t y = mt[(*idx)++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  // random.choice(seq) index: _randbelow_with_getrandbits(n), n < 2^32
  uint32_t randbelow(uint32_t n) {
    int k = 0;
    for (uint32_t v = n; v; v >>= 1) ++k;  // n.bit_length()
    uint32_t r = genrand() >> (32 - k);
    while (r >= n) r = genrand() >> (32 - k);
    return r;
  }
};

// ------------------------------------------------------------------------------------------
// CPython: set of (y, x) tuples
// ------------------------------------------------------------------------------------------
struct Cell {
  int32_t y, x;
  bool operator==(const Cell& o) const { return y == o.y && x == o.x; }
};

uint64_t hash_int(int64_t v) {
  const int64_t mod = ((int64_t)1 << 61) - 1;
  int64_t h = v >= 0 ? v % mod : -((-v) % mod);
  if (h == -1) h = -2;
  return (uint64_t)h;
}
uint64_t hash_cell(Cell c) {  // tuplehash of a 2-tuple of ints (CPython >= 3.8)
  const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P5 = 2870177450012600261ull;
  uint64_t acc = P5;
  const uint64_t lanes[2] = {hash_int(c.y), hash_int(c.x)};
  for (uint64_t lane : lanes) {
    acc += lane * P2;
    acc = (acc << 31) | (acc >> 33);
    acc *= P1;
  }
  acc += 2ull ^ (P5 ^ 3527539ull);
  if (acc == ~0ull) return 1546275796ull;
  return acc;
}

// Bump allocator for the per-episode sets: tables are carved out of a few large blocks and all
// released at once when the next episode starts (a plan of 256 episodes used to cost ~4000 mallocs).
struct Arena {
  std::vector<std::unique_ptr<uint8_t[]>> blocks;
  std::vector<size_t> sizes;
  size_t block = 0, off = 0;
  void reset() { block = 0; off = 0; }
  void* alloc(size_t bytes) {
    bytes = (bytes + 15) & ~size_t(15);
    for (;; ++block, off = 0) {
      if (block == blocks.size()) {
        const size_t cap = std::max<size_t>(bytes, size_t(1) << 18);
        blocks.emplace_back(new uint8_t[cap]);
        sizes.push_back(cap);
      }
      if (off + bytes <= sizes[block]) {
        void* p = blocks[block].get() + off;
        off += bytes;
        return p;
      }
    }
  }
};

// CPython's set: open addressing, LINEAR_PROBES = 9, perturb shift 5, dummy entries, resize rules of
// set_add_entry / set_merge.  Tables live in an Arena; copying a PySet aliases its table (copies are
// only ever read).
struct PySet {
  enum : uint8_t { kEmpty = 0, kActive = 1, kDummy = 2 };
  struct Entry {
    uint64_t hash;
    Cell key;
    uint8_t state;
  };
  Arena* arena;
  Entry* table;
  size_t mask = 7, fill = 0, used = 0;
  static constexpr size_t LP = 9;

  static Entry* fresh(Arena* a, size_t n) {
    Entry* t = static_cast<Entry*>(a->alloc(n * sizeof(Entry)));
    for (size_t i = 0; i < n; ++i) t[i] = Entry{0, {0, 0}, kEmpty};
    return t;
  }
  explicit PySet(Arena* a) : arena(a), table(fresh(a, 8)) {}

  static void insert_clean(Entry* t, size_t mask, Cell key, uint64_t h) {
    size_t perturb = h, i = h & mask;
    for (;;) {
      if (t[i].state == kEmpty) { t[i] = Entry{h, key, kActive}; return; }
      if (i + LP <= mask)
        for (size_t j = 1; j <= LP; ++j)
          if (t[i + j].state == kEmpty) { t[i + j] = Entry{h, key, kActive}; return; }
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  void resize(size_t minused) {
    size_t newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    const Entry* old = table;
    const size_t old_size = mask + 1;
    table = fresh(arena, newsize);
    mask = newsize - 1;
    fill = used;
    for (size_t i = 0; i < old_size; ++i)
      if (old[i].state == kActive) insert_clean(table, mask, old[i].key, old[i].hash);
  }
  void add(Cell key, uint64_t h) {
    size_t i = h & mask, perturb = h;
    long free_slot = -1;
    for (;;) {
      size_t probes = (i + LP <= mask) ? LP : 0, j = i;
      for (;;) {
        Entry& e = table[j];
        if (e.state == kEmpty) {
          if (free_slot >= 0) {
            used++;
            table[(size_t)free_slot] = Entry{h, key, kActive};
            return;
          }
          fill++;
          used++;
          e = Entry{h, key, kActive};
          if (fill * 5 < mask * 3) return;
          resize(used > 50000 ? used * 2 : used * 4);
          return;
        }
        if (e.state == kDummy) free_slot = (long)j;
        else if (e.hash == h && e.key == key) return;
        ++j;
        if (probes == 0) break;
        --probes;
      }
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  void add(Cell key) { add(key, hash_cell(key)); }
  long find(Cell key) const {
    const uint64_t h = hash_cell(key);
    size_t i = h & mask, perturb = h;
    for (;;) {
      size_t probes = (i + LP <= mask) ? LP : 0, j = i;
      for (;;) {
        const Entry& e = table[j];
        if (e.state == kEmpty) return -1;
        if (e.state == kActive && e.hash == h && e.key == key) return (long)j;
        ++j;
        if (probes == 0) break;
        --probes;
      }
      perturb >>= 5;
      i = (i * 5 + 1 + perturb) & mask;
    }
  }
  bool contains(Cell key) const { return find(key) >= 0; }
  bool remove(Cell key) {
    const long j = find(key);
    if (j < 0) return false;
    table[(size_t)j].state = kDummy;
    used--;
    return true;
  }
  // set_merge: `so |= other`, also the body of copying / union
  void merge(const PySet& o) {
    if (o.table == table || o.used == 0) return;
    if ((fill + o.used) * 5 >= mask * 3) resize((used + o.used) * 2);
    if (fill == 0 && mask == o.mask && o.fill == o.used) {
      for (size_t i = 0; i <= o.mask; ++i)
        if (o.table[i].state == kActive) table[i] = o.table[i];
      fill = o.fill;
      used = o.used;
      return;
    }
    if (fill == 0) {
      fill = used = o.used;
      for (size_t i = 0; i <= o.mask; ++i)
        if (o.table[i].state == kActive) insert_clean(table, mask, o.table[i].key, o.table[i].hash);
      return;
    }
    for (size_t i = 0; i <= o.mask; ++i)
      if (o.table[i].state == kActive) add(o.table[i].key, o.table[i].hash);
  }
  template <typename F>
  void for_each(F f) const {
    for (size_t i = 0; i <= mask; ++i)
      if (table[i].state == kActive) f(table[i].key);
  }
};

int64_t floordiv64(int64_t a, int64_t b) {
  int64_t q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
int64_t pymod(int64_t a, int64_t b) {
  int64_t r = a % b;
  return r < 0 ? r + b : r;
}

// bbox_positions (simple_env.py:270-321), with the reference's set construction sequence
PySet box_cells(Arena* arena, const int64_t* b, int P, int rows, int cols) {
  const int64_t x1 = b[0], y1 = b[1], x2 = b[2], y2 = b[3];
  const int64_t py_lo = floordiv64(y1, P), py_hi = floordiv64(y2, P), px_lo = floordiv64(x1, P), px_hi = floordiv64(x2, P);
  PySet cells(arena);
  for (int64_t y = py_lo; y <= py_hi; ++y)
    for (int64_t x = px_lo; x <= px_hi; ++x) {
      const int64_t oh = std::min<int64_t>((y + 1) * P, y2) - std::max<int64_t>(y * P, y1);
      const int64_t ow = std::min<int64_t>((x + 1) * P, x2) - std::max<int64_t>(x * P, x1);
      if ((double)(oh * ow) / (double)((int64_t)P * P) > 0.05) cells.add(Cell{(int32_t)y, (int32_t)x});
    }
  cells.add(Cell{(int32_t)floordiv64(floordiv64(y1 + y2, 2), P), (int32_t)floordiv64(floordiv64(x1 + x2, 2), P)});
  PySet in_x(arena);
  cells.for_each([&](Cell c) { if (c.x >= 0 && c.x < cols) in_x.add(c); });
  PySet in_y(arena);
  in_x.for_each([&](Cell c) { if (c.y >= 0 && c.y < rows) in_y.add(c); });
  return in_y;
}

}  // namespace

struct jn_plan {
  std::vector<int32_t> start, seg_begin, seg_to, seg_tgt, draw_begin, det_begin, det_yx, final_pos;
  std::vector<uint8_t> seg_flags, draws;
  char error[256] = "";
  // scratch reused from episode to episode (and from call to call)
  Arena arena;
  std::vector<uint8_t> is_box;
  std::vector<Cell> visited, empties, keypoints, ties;
  std::vector<int64_t> slots;
  std::vector<PySet> cells;
  // jn_plan_start / jn_plan_wait
  std::thread worker;
  int worker_status = JN_OK;
};

namespace {

struct Episode {
  jn_plan& out;
  NumpyRng rng;
  PyRandom& py;
  int rows, cols;
  std::vector<uint8_t>& is_box;  // bbox_patches membership (order never matters for it)
  std::vector<Cell>& visited;    // visited_bbox_patches (membership / removal only)
  Cell pos{0, 0};

  bool inside(Cell c) const { return is_box[(size_t)c.y * cols + c.x] != 0; }
  void place(Cell c) {  // reset(position) with visited=None (simple_env.py:335-343)
    pos = c;
    visited.clear();
    if (inside(c)) visited.push_back(c);
  }
  void draw_move() { out.draws.push_back((uint8_t)rng.integers(0, 8)); }
  void walk(Cell to, Cell tgt, int first) {  // visit_point (simple_env.py:631-664)
    out.seg_to.push_back(to.y); out.seg_to.push_back(to.x);
    out.seg_tgt.push_back(tgt.y); out.seg_tgt.push_back(tgt.x);
    out.seg_flags.push_back((uint8_t)first);
    int y = pos.y, x = pos.x;
    while (y != to.y || x != to.x) {
      y += (to.y > y) - (to.y < y);
      x += (to.x > x) - (to.x < x);
      if (y == tgt.y && x == tgt.x) draw_move();
    }
    place(Cell{y, x});
  }
};

}  // namespace

extern "C" {

int jn_plan_create(jn_plan** out) {
  if (!out) return JN_ERR_INVALID;
  *out = new jn_plan();
  return JN_OK;
}
void jn_plan_destroy(jn_plan* p) {
  if (p && p->worker.joinable()) p->worker.join();
  delete p;
}
const char* jn_plan_error(const jn_plan* p) { return p ? p->error : "null plan"; }

int jn_plan_run(jn_plan* plan, int n, const int64_t* boxes, const int32_t* n_boxes, int max_boxes,
                const int32_t* rows, const int32_t* cols, int patch_size, const uint64_t* seeds,
                const uint8_t* has_seed, int min_keypoints, int max_keypoints, int binomial,
                const int32_t* start_yx, uint32_t* mt_state) {
  if (!plan || n < 0 || !rows || !cols || !mt_state || patch_size < 1 || (max_boxes > 0 && (!boxes || !n_boxes)))
    return JN_ERR_INVALID;
  jn_plan& o = *plan;
  o.start.clear(); o.seg_to.clear(); o.seg_tgt.clear(); o.seg_flags.clear(); o.draws.clear(); o.det_yx.clear();
  o.final_pos.clear();
  o.seg_begin.assign(1, 0); o.draw_begin.assign(1, 0); o.det_begin.assign(1, 0);
  PyRandom py{mt_state, mt_state + 624};
  std::unique_ptr<std::random_device> entropy;  // only opened for unseeded episodes
  for (int e = 0; e < n; ++e) {
    const int R = rows[e], C = cols[e], nb = max_boxes > 0 ? n_boxes[e] : 0;
    if (R < 1 || C < 1 || R > 4096 || C > 4096) {
      snprintf(o.error, sizeof(o.error), "episode %d: grid %dx%d outside the native planner's range (1..4096)", e, R, C);
      return JN_ERR_UNSUPPORTED;
    }
    o.arena.reset();
    o.is_box.assign((size_t)R * C, 0);
    o.visited.clear();
    Episode ep{o, NumpyRng(), py, R, C, o.is_box, o.visited};
    if (has_seed && has_seed[e]) ep.rng.seed_u64(seeds[e]);
    else {
      if (!entropy) entropy.reset(new std::random_device());
      uint32_t w[4] = {(*entropy)(), (*entropy)(), (*entropy)(), (*entropy)()};
      ep.rng.seed(w, 4);
    }
    const int64_t* eb = boxes + (size_t)e * max_boxes * 4;
    // bbox_positions of every box: the reference recomputes them three times (env construction,
    // init_sample, build_keypoints_trajectory); they are pure functions of the box, so once is enough
    std::vector<PySet>& cells = o.cells;
    cells.clear();
    for (int k = 0; k < nb; ++k) cells.push_back(box_cells(&o.arena, eb + 4 * k, patch_size, R, C));
    // env construction: bbox_patches (membership only)
    for (int k = 0; k < nb; ++k) cells[(size_t)k].for_each([&](Cell c) { ep.is_box[(size_t)c.y * C + c.x] = 1; });
    // init_sample: detection patches = every box patch + one random empty patch, in set order
    PySet det(&o.arena);
    for (int k = 0; k < nb; ++k) cells[(size_t)k].for_each([&](Cell c) { det.add(c); });
    std::vector<Cell>& empties = o.empties;
    empties.clear();
    for (int y = 0; y < R; ++y)
      for (int x = 0; x < C; ++x)
        if (!det.contains(Cell{y, x})) empties.push_back(Cell{y, x});
    if (!empties.empty()) det.add(empties[(size_t)ep.rng.integers(0, (int64_t)empties.size())]);
    det.for_each([&](Cell c) { o.det_yx.push_back(c.y); o.det_yx.push_back(c.x); });
    o.det_begin.push_back((int32_t)(o.det_yx.size() / 2));
    // start position: given, or y then x from the numpy stream
    Cell start;
    if (start_yx) start = Cell{start_yx[2 * e], start_yx[2 * e + 1]};
    else {
      start.y = (int32_t)ep.rng.integers(0, R);
      start.x = (int32_t)ep.rng.integers(0, C);
    }
    if (start.y < 0 || start.y >= R || start.x < 0 || start.x >= C) {
      snprintf(o.error, sizeof(o.error), "episode %d: start position (%d, %d) outside the %dx%d grid", e, start.y, start.x, R, C);
      return JN_ERR_INVALID;
    }
    ep.place(start);
    o.start.push_back(start.y); o.start.push_back(start.x);
    // build_keypoints_trajectory: greedy L1-nearest, ties through random.choice in set order
    PySet todo(&o.arena);
    for (int k = 0; k < nb; ++k) todo.merge(cells[(size_t)k]);
    for (Cell v : ep.visited) todo.remove(v);
    std::vector<Cell>&keypoints = o.keypoints, &ties = o.ties;
    keypoints.clear();
    Cell here = ep.pos;
    while (todo.used > 0) {
      long best = -1;
      ties.clear();
      todo.for_each([&](Cell c) {
        const long d = std::labs((long)c.x - here.x) + std::labs((long)c.y - here.y);
        if (best < 0 || d < best) { best = d; ties.clear(); }
        if (d == best) ties.push_back(c);
      });
      here = ties[py.randbelow((uint32_t)ties.size())];
      keypoints.push_back(here);
      todo.remove(here);
    }
    if (keypoints.empty()) {
      Cell k;
      k.y = (int32_t)ep.rng.integers(0, R);
      k.x = (int32_t)ep.rng.integers(0, C);
      keypoints.push_back(k);
    }
    // random key points: how many, and before which key point
    const int64_t n_random = ep.rng.integers(min_keypoints, (int64_t)max_keypoints + 1);
    std::vector<int64_t>& slots = o.slots;
    slots.clear();
    for (int64_t i = 0; i < n_random; ++i) slots.push_back(ep.rng.integers(0, (int64_t)keypoints.size()));
    for (size_t k = 0; k < keypoints.size(); ++k) {
      const Cell kp = keypoints[k];
      int first = 1;
      if (ep.pos == kp) ep.draw_move();  // the opening best action would be STOP
      for (;;) {
        size_t hit = slots.size();
        for (size_t i = 0; i < slots.size(); ++i)
          if (slots[i] == (int64_t)k) { hit = i; break; }
        if (hit == slots.size()) break;
        Cell detour;
        if (binomial) {  // x is drawn first (simple_env.py:705-707)
          const int64_t dx = ep.rng.binomial(C, 0.5) - C / 2;
          const int64_t dy = ep.rng.binomial(R, 0.5) - R / 2;
          detour = Cell{(int32_t)pymod(kp.y + dy, R), (int32_t)pymod(kp.x + dx, C)};
        } else {
          detour.y = (int32_t)ep.rng.integers(0, R);
          detour.x = (int32_t)ep.rng.integers(0, C);
        }
        ep.walk(detour, kp, first);
        first = 0;
        slots.erase(slots.begin() + (long)hit);
      }
      ep.walk(kp, kp, first);
    }
    o.seg_begin.push_back((int32_t)o.seg_flags.size());
    o.draw_begin.push_back((int32_t)o.draws.size());
    o.final_pos.push_back(ep.pos.y); o.final_pos.push_back(ep.pos.x);
  }
  return JN_OK;
}

int jn_plan_start(jn_plan* plan, int n, const int64_t* boxes, const int32_t* n_boxes, int max_boxes,
                  const int32_t* rows, const int32_t* cols, int patch_size, const uint64_t* seeds,
                  const uint8_t* has_seed, int min_keypoints, int max_keypoints, int binomial,
                  const int32_t* start_yx, uint32_t* mt_state) {
  if (!plan || plan->worker.joinable()) return JN_ERR_INVALID;  // one run at a time
  plan->worker_status = JN_OK;
  try {
    plan->worker = std::thread([=] {
      plan->worker_status = jn_plan_run(plan, n, boxes, n_boxes, max_boxes, rows, cols, patch_size, seeds, has_seed,
                                        min_keypoints, max_keypoints, binomial, start_yx, mt_state);
    });
  } catch (...) {  // no thread to be had: plan on the caller's
    plan->worker_status = jn_plan_run(plan, n, boxes, n_boxes, max_boxes, rows, cols, patch_size, seeds, has_seed,
                                      min_keypoints, max_keypoints, binomial, start_yx, mt_state);
  }
  return JN_OK;
}

int jn_plan_wait(jn_plan* plan) {
  if (!plan) return JN_ERR_INVALID;
  if (plan->worker.joinable()) plan->worker.join();
  return plan->worker_status;
}

int jn_plan_sizes(const jn_plan* p, int* n_segments, int* n_draws, int* n_det) {
  if (!p) return JN_ERR_INVALID;
  if (n_segments) *n_segments = (int)p->seg_flags.size();
  if (n_draws) *n_draws = (int)p->draws.size();
  if (n_det) *n_det = (int)(p->det_yx.size() / 2);
  return JN_OK;
}

// Copies the plan into caller buffers: start [n,2], seg_begin [n+1], seg_to [S,2], seg_tgt [S,2],
// draw_begin [n+1], det_begin [n+1], det_yx [D,2] (all int32), seg_flags [S], draws [Q] (uint8).
int jn_plan_export(const jn_plan* p, int32_t* start, int32_t* seg_begin, int32_t* seg_to, int32_t* seg_tgt,
                   int32_t* draw_begin, int32_t* det_begin, int32_t* det_yx, uint8_t* seg_flags, uint8_t* draws) {
  if (!p) return JN_ERR_INVALID;
  auto put32 = [](int32_t* dst, const std::vector<int32_t>& v) { if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * 4); };
  put32(start, p->start); put32(seg_begin, p->seg_begin); put32(seg_to, p->seg_to); put32(seg_tgt, p->seg_tgt);
  put32(draw_begin, p->draw_begin); put32(det_begin, p->det_begin); put32(det_yx, p->det_yx);
  if (seg_flags && !p->seg_flags.empty()) memcpy(seg_flags, p->seg_flags.data(), p->seg_flags.size());
  if (draws && !p->draws.empty()) memcpy(draws, p->draws.data(), p->draws.size());
  return JN_OK;
}

}  // extern "C"
