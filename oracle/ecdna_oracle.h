/*
 * ecdna_oracle.h -- CPU ORACLE for the ecDNA SSA hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under ecdna-evo_b200/ may include, link or call this.  It is used by
 * tests/, by __graft_entry__.smoke() as the checker, and by bench.py's
 * cpu_baseline / --impl reference legs as the timed CPU implementation.
 *
 * PARITY UNPINNED.  The arithmetic of the reference's hot path lives in crates
 * that are not under /root/reference and not in this image: sosa 3.0.3 (event
 * loop), ecdna-lib 3.0.2 (state container), rand_distr 0.4.3 (Binomial, Exp),
 * rand 0.8.5 / rand_chacha 0.3.1 (ChaCha8, integer/float conversions); see
 * Cargo.lock:423-426,802-836,944-947.  No Rust toolchain exists here, and the
 * reference's own tests hold no numeric golden vector for this path (only the
 * invariants of proliferation.rs:159-286 and segregation.rs:223-291, which
 * tests/test_oracle_reference_props.py re-runs against this file).  What is
 * restated from the reference's files is cited as file:line; what is restated
 * from the published algorithms of the absent crates is marked [RECALL].
 *
 * Two state layouts, one event semantics:
 *   state 0 "vector"    the reference's layout: nminus counter + one u16 per
 *                       ecDNA+ cell, uniform index pick + swap-remove, daughters
 *                       pushed k1 then k2 (memory.md:5-8, proliferation.rs:57,85-88,109)
 *   state 1 "histogram" the GPU's layout: count of cells per copy number, the
 *                       uniform draw mapped to a class in the canonical order
 *                       (k mod 32, k).  This is the bit-exact specification of the
 *                       CUDA kernel's native mode.
 * Three random sources:
 *   rng 0 "rand"    ChaCha8(seed).set_stream(idx) (main.rs:57-58) with rand-0.8
 *                   conversions, ziggurat Exp1, BINV/BTPE binomial [RECALL]
 *   rng 1 "philox"  Philox4x32-10, key=(seed), counter=(event, slot, run): the
 *                   GPU's native stream v2 (slot layout in ecdna_oracle.cpp PhiloxSource): Gillespie's
 *                   direct method - one exponential waiting time with the summed propensity (a deterministic
 *                   f32 log) and one uniform that picks the reaction - and Binomial(2k,1/2) as the
 *                   popcount of 2k random bits (exact)
 *   rng 2 "replay"  consumes a decision stream {event, dt, k, k1} (histogram state
 *                   only); the stream is what either state emits as trace_out.
 */
#ifndef ECDNA_ORACLE_H
#define ECDNA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* process.rs:20-29 (the four live variants) in sosa reaction order main.rs:140-145 */
enum { ORC_EV_BIRTH_NMINUS = 0, ORC_EV_BIRTH_NPLUS = 1, ORC_EV_DEATH_NMINUS = 2, ORC_EV_DEATH_NPLUS = 3 };
/* clap_app.rs:232-238 */
enum { ORC_SEG_DETERMINISTIC = 0, ORC_SEG_BINOMIAL_NO_UNEVEN = 1, ORC_SEG_BINOMIAL = 2, ORC_SEG_BINOMIAL_NO_NMINUS = 3 };
/* sosa::StopReason [RECALL] + the two conditions the reference turns into panics */
enum {
  ORC_STOP_NO_INDIVIDUALS = 0, ORC_STOP_MAX_ITERS = 1, ORC_STOP_MAX_TIME = 2, ORC_STOP_MAX_CELLS = 3,
  ORC_STOP_ABSORBING = 4,      /* cells left but every propensity is zero */
  ORC_STOP_COPY_OVERFLOW = 5,  /* k*2 overflows u16: proliferation.rs:63-67 panics */
  ORC_STOP_HIST_OVERFLOW = 6,  /* copy number >= hist_cap of the caller's buffer */
  ORC_STOP_REPLAY_END = 7,     /* replay stream exhausted */
  ORC_STOP_REPLAY_BAD = 8      /* replay record inconsistent with the state */
};

typedef struct {
  float dt;        /* waiting time added to the clock, f32 (process.rs:184) */
  uint16_t k;      /* copies of the cell picked (ecDNA+ birth or death), else 0 */
  uint16_t k1;     /* first daughter's copies (ecDNA+ birth), else 0 */
  uint8_t event;   /* ORC_EV_* */
  uint8_t pad[3];
} orc_replay_event;

typedef struct {
  float b0, b1, d0, d1;      /* main.rs:31-34 */
  uint32_t segregation;      /* ORC_SEG_* */
  uint32_t state;            /* 0 vector, 1 histogram */
  uint32_t rng;              /* 0 rand, 1 philox, 2 replay */
  uint32_t bd_count_mode;    /* 0: stop on true cells; 1: stop on sosa's population sum
                                (birth-death state is [n-,n+,n-,n+], process.rs:339-344) */
  uint64_t max_cells;        /* clap_app.rs:208 */
  uint64_t max_iter;         /* main.rs:23 */
  float max_time;            /* clap_app.rs:205 */
  uint32_t birth_death;      /* clap_app.rs:194-200 (only matters for bd_count_mode 1) */
  uint64_t seed;             /* clap_app.rs:63-64 */
  uint64_t run_idx;          /* main.rs:56, 214-215 */
  uint32_t n_init;           /* initial distribution, sparse; k==0 entry is nminus */
  const uint16_t* init_k;
  const uint64_t* init_c;
  uint32_t n_snap;           /* clap_app.rs:102-134, sorted ascending */
  const uint64_t* snap_cells;
  uint32_t dyn_points;       /* dynamics: CHANGELOG.md:34-40 (300 bins, t<=30) */
  float dyn_dt;
  const orc_replay_event* replay_in;
  uint64_t replay_len;
  /* vector state + rng "rand": when set, these u64 (the log of a recording RngCore around the REAL
     reference's ChaCha8Rng) replace the generator; u64_len of orc_out then
     reports how many were consumed and a stream that ends early stops with ORC_STOP_REPLAY_END */
  const uint64_t* u64_in;
  uint64_t u64_in_len;
} orc_opts;

typedef struct {
  uint32_t stop_reason;
  uint32_t kmax;             /* largest copy number ever present */
  uint64_t nminus, nplus, n_events;
  float time;
  uint32_t n_snap_taken;
  uint64_t hash;             /* sum_k c_k*w(k) mod 2^64 of the final histogram */
  uint64_t chain;            /* chained per-event digest of (hash, nminus, time bits) */
  uint64_t sum_k;            /* sum over ecDNA+ events of (kmax+1): flat-model read bytes/4 */
  uint64_t n_div, n_death;   /* ecDNA+ divisions, ecDNA+ deaths */
  uint32_t dyn_count;
  /* caller-owned buffers (may be NULL / 0) */
  uint32_t hist_cap;
  uint64_t* hist;            /* dense final histogram, hist[0] = nminus */
  orc_replay_event* trace_out;
  uint64_t trace_cap, trace_len;
  uint64_t* traj_out;        /* per event: nminus, nplus, time bits, hash (4 x u64) */
  uint64_t traj_cap;
  uint64_t* snap_hist;       /* [n_snap][hist_cap] */
  uint64_t* snap_cells_out;  /* [n_snap] */
  float* snap_time;          /* [n_snap] */
  float* dyn_out;            /* [dyn_points][5]: nminus, nplus, mean, variance, entropy */
  uint64_t* u64_out;         /* vector state + rng "rand": every u64 the generator handed out, in order */
  uint64_t u64_cap, u64_len;
} orc_out;

int orc_run(const orc_opts* o, orc_out* out);

/* many replicates over host threads with a dynamic queue (rayon par_iter, main.rs:221-224).
   Returns total events; per-run outputs optional (arrays of n_runs). */
uint64_t orc_run_batch(const orc_opts* o, uint64_t idx_begin, uint64_t n_runs, int n_threads,
                       uint64_t* nminus, uint64_t* nplus, float* time, uint64_t* n_events,
                       uint32_t* stop, uint64_t* hist /* [n_runs][hist_cap] or NULL */, uint32_t hist_cap,
                       const float* rates_per_run /* [n_runs][4] or NULL */);

/* ABC over a batch of prior draws (abc.md:38-55): per draw one replicate with rates_per_run[i] and the four
   distances of its final distribution to `target` (dense, [0] = cells without ecDNA): KS, relative mean,
   relative entropy, relative frequency [RECALL R8]; accept = every distance <= its threshold (a negative
   threshold is ignored).  Returns total events. */
uint64_t orc_abc_batch(const orc_opts* o, uint64_t idx_begin, uint64_t n_runs, int n_threads, const float* rates_per_run,
                       const uint64_t* target, uint32_t target_len, const float thresholds[4], uint32_t hist_cap,
                       float* distances, uint8_t* accept, uint64_t* n_events, uint32_t* stop);

/* summary statistics over a dense histogram (hist[0] = cells without ecDNA) [RECALL R8] */
void orc_stats(const uint64_t* hist, uint32_t cap, float* mean, float* frequency, float* entropy, float* variance);
float orc_ks_distance(const uint64_t* h1, uint32_t cap1, const uint64_t* h2, uint32_t cap2);

/* building blocks exposed for known-answer and distribution tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_chacha8_u64(uint64_t seed, uint64_t stream, uint64_t n, uint64_t* out);
void orc_chacha_block(const uint32_t key[8], uint64_t counter, uint64_t stream, int rounds, uint32_t out[16]);
void orc_seed_from_u64(uint64_t seed, uint32_t key[8]);
float orc_neg_log_u24(uint32_t m);
uint32_t orc_binomial_half_philox(uint64_t seed, uint64_t run, uint32_t event, uint32_t attempt, uint32_t n);
uint64_t orc_pick_philox(uint64_t seed, uint64_t run, uint32_t event, uint64_t n);
void orc_rand_binomial(uint64_t seed, uint64_t stream, uint64_t n, double p, uint64_t count, uint64_t* out);
void orc_rand_exp1_f32(uint64_t seed, uint64_t stream, uint64_t count, float* out);
void orc_rand_gen_range(uint64_t seed, uint64_t stream, uint64_t n, uint64_t count, uint64_t* out);
/* one division / death applied to a vector-state distribution; used to replay the reference's
   property tests (proliferation.rs:159-286).  hist is dense, modified in place. */
int orc_apply_event(uint64_t* hist, uint32_t cap, uint32_t event, uint32_t segregation, uint64_t seed,
                    uint32_t* k_out, uint32_t* k1_out, uint32_t* k2_out, uint32_t* uneven_out);
int orc_segregate(uint32_t rule, uint32_t copies, uint64_t seed, uint64_t* k1, uint64_t* k2, uint32_t* uneven);
uint64_t orc_hist_weight(uint32_t k);
/* EcDNADistribution::into_subsampled (src/main.rs:110-123) in the library's native mode: `want` cells drawn
   without replacement from hist[0..len) ([0] = cells without ecDNA), one cell per draw, Philox counter
   (draw, 0x40000000 + j + 65536*(attempt/2), run_lo, run_hi); see csrc/subsample.cuh.  out[0..len). */
void orc_subsample(const uint64_t* hist, uint32_t len, uint64_t want, uint64_t seed, uint64_t run_idx, uint32_t j,
                   uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif
