// Drop-in `class Mandelbrot` (include/newman_b200/mandelbrot.h) over the device C-ABI.
// Mirrors the reference's public behaviour (reference mandelbrot.cpp) for the view state, precision
// policy, view transforms and multisample rescale; replaces the per-pixel work with one GPU frame.
#include "../../include/newman_b200/mandelbrot.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>

#include "../../include/newman_b200.h"
#include "hp_host.h"
#include "multi_host.h"

#include <atomic>
#include <cstring>
#include <memory>
#include <thread>

namespace newman_b200 {

struct Cancelled : public std::runtime_error {   // thrown through a frame that Mandelbrot::cancel() abandoned
  Cancelled() : std::runtime_error("newman_b200: frame cancelled") {}
};

class Engine {
public:
  nm_ctx* ctx = nullptr;
  int device;
  bool owned = true;
  std::atomic<bool> cancel_requested{false};   // set by cancel() (any thread), cleared when the next frame starts
  explicit Engine(int dev) : device(dev) {
    int rc = nm_create(dev, &ctx);
    if (rc != NM_OK) throw std::runtime_error(std::string("newman_b200: ") + nm_last_error(nullptr));
  }
  explicit Engine(nm_ctx* borrowed) : ctx(borrowed), device(-1), owned(false) {}   // a render group's rank context
  ~Engine() { if (ctx && owned) nm_destroy(ctx); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  void check(int rc, const char* what) {
    if (rc == NM_ECANCELLED || cancel_requested.load()) throw Cancelled();
    if (rc != NM_OK) throw std::runtime_error(std::string("newman_b200: ") + what + ": " + nm_last_error(ctx));
  }
};

}  // namespace newman_b200

using newman_b200::DeepTablesHost;
using newman_b200::ViewHP;

static_assert(sizeof(RenderGrid::EscapeValue) == sizeof(nm_escape), "EscapeValue must be the 8-byte device record");

struct Mandelbrot::Signature {
  int N, nr, nc, max_secondary;
  double tol, gtol;
  mpf_class cre, cim, sre, sim;
};

namespace {
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
mp_bitcnt_t max_prec(const HPComplex& a, const HPComplex& b) {
  mp_bitcnt_t p = a.re.get_prec();
  if (a.im.get_prec() > p) p = a.im.get_prec();
  if (b.re.get_prec() > p) p = b.re.get_prec();
  if (b.im.get_prec() > p) p = b.im.get_prec();
  return p;
}

// NM_DEBUG_HOST=1: where the host time of a frame goes (stderr)
struct HostTrace {
  bool on; double t0; const char* what;
  explicit HostTrace(const char* w) : on(getenv("NM_DEBUG_HOST") != nullptr), t0(now_s()), what(w) {}
  void lap(const char* stage) {
    if (!on) return;
    const double t = now_s();
    fprintf(stderr, "nm host %-18s %-28s %8.1f ms\n", what, stage, 1e3 * (t - t0));
    t0 = t;
  }
};

struct RoundParams {
  int N, max_secondary, force_floatexp;
  double tol, gtol;
  int host_threads;   // build_tables: 4 or more (0 = all cores) pipelines orbit and series
  bool orbit_trunc;   // exact mode's probe rendering: K3 iterates against the truncated orbit (NM_TABLES_ORBIT_TRUNCATED)
};

// One reference after another until no sample is left glitched: the frame (or the listed samples)
// against T, then the flagged samples against a secondary reference = the flagged sample with the
// earliest flag iteration (lowest id on ties), ...; after max_secondary such rounds the rest is
// finished in one rebasing pass. T is overwritten by the secondary tables.
void run_rounds(newman_b200::Engine& eng, const ViewHP& v, const RoundParams& rp, DeepTablesHost& T, int cmode,
                const std::vector<uint8_t>& mask, const std::vector<int32_t>* first_list, newman_b200::FrameInfo& info,
                std::vector<int32_t>* primary_glitched = nullptr) {
  nm_ctx* ctx = eng.ctx;
  std::vector<int32_t> rq_pix, rq_iter;
  if (first_list) rq_pix = *first_list;
  int round = 0;
  for (;;) {
    // Coefficients beyond double range (pixel pitch < ~1e-97, where the reference dies with SIGFPE):
    // hand them over as mantissa + exponent and let K2 evaluate the series in floatexp (level 1).
    // Below a pitch of 2^-380 (~4e-115) delta*delta, and later delta and eps themselves, leave double
    // range too: eps goes over as mantissa + exponent and K3 iterates scaled states (level 2).
    int fe = T.finite ? 0 : 1;
    if (T.pitch_exp < -380) fe = 2;
    if (rp.force_floatexp > fe) fe = rp.force_floatexp > 2 ? 2 : rp.force_floatexp;
    nm_deep_tables t;
    memset(&t, 0, sizeof t);
    t.M = T.M; t.N = rp.N; t.has_escape = T.has_escape ? 1 : 0; t.flags = rp.orbit_trunc ? NM_TABLES_ORBIT_TRUNCATED : 0;
    t.tol = rp.tol; t.glitch_tol = rp.gtol;
    t.x_hi = T.x_hi.data(); t.x_lo = T.x_lo.data();
    t.a = fe ? T.a_m.data() : T.a.data(); t.b = fe ? T.b_m.data() : T.b.data(); t.c = fe ? T.c_m.data() : T.c.data();
    t.a_exp = fe ? T.a_e.data() : nullptr; t.b_exp = fe ? T.b_e.data() : nullptr; t.c_exp = fe ? T.c_e.data() : nullptr;
    t.eps_re_exp = fe == 2 ? T.eps_re_e.data() : nullptr; t.eps_im_exp = fe == 2 ? T.eps_im_e.data() : nullptr;
    info.floatexp = fe;
    const bool last = round >= rp.max_secondary;
    const bool listed = round > 0 || first_list;
    eng.check(nm_frame_deep(ctx, &t, fe == 2 ? T.eps_re_m.data() : T.eps_re.data(), v.nc,
                            fe == 2 ? T.eps_im_m.data() : T.eps_im.data(), v.nr, cmode,
                            cmode == NM_CARDIOID_MASK ? mask.data() : nullptr, listed ? rq_pix.data() : nullptr,
                            listed ? (int64_t)rq_pix.size() : 0, last ? NM_MODE_REBASE : NM_MODE_REQUEUE),
              "nm_frame_deep");
    eng.check(nm_launch(ctx), "nm_launch");
    nm_stats st; eng.check(nm_frame_stats(ctx, &st), "nm_frame_stats");
    info.executed_iters += st.executed_iters; info.series_evals += st.series_evals;
    info.skipped_pixels += st.skipped_pixels; info.rebased += st.rebased; info.fixups += st.fixups;
    info.kernel_launches += st.kernel_launches;
    info.device_ms += st.ms_k1 + st.ms_k2 + st.ms_k3;
    info.references++;
    int64_t n_rq = nm_frame_requeue(ctx, nullptr, nullptr, 0);
    if (n_rq < 0) eng.check((int)n_rq, "nm_frame_requeue");
    if (n_rq == 0) break;
    info.glitched += (unsigned long long)n_rq;
    rq_pix.resize((size_t)n_rq); rq_iter.resize((size_t)n_rq);
    nm_frame_requeue(ctx, rq_pix.data(), rq_iter.data(), n_rq);
    if (round == 0 && primary_glitched) *primary_glitched = rq_pix;
    // next reference: the glitched sample flagged earliest, lowest pixel id on ties
    size_t best = 0;
    for (size_t i = 1; i < rq_pix.size(); i++)
      if (rq_iter[i] < rq_iter[best] || (rq_iter[i] == rq_iter[best] && rq_pix[i] < rq_pix[best])) best = i;
    const double t_hp = now_s();
    newman_b200::build_tables(v, rq_pix[best] / v.nc, rq_pix[best] % v.nc, T, rp.host_threads);
    info.host_precompute_s += now_s() - t_hp;
    cmode = NM_CARDIOID_NONE;  // listed samples already passed the cardioid test
    round++;
  }
}

// findProbe (mandelbrot.cpp:73-95) with the GPU doing the bulk of it. The reference computes one full
// arbitrary-precision orbit per candidate (3*ceil(nc/2) + ceil(nr/2) of them: 27 360 x up to N
// iterations for a 15 360 x 8 640 grid) to keep "the first one in scan order with the longest orbit".
// A candidate's orbit length IS its escape count, so: one arbitrary-precision reference (the centre
// sample), all candidates rendered against it by K2/K3 like any other samples (secondary references
// for glitched ones), then only the short-list — the `kShort` longest by that count and everything
// within 0.2 % of the maximum, plus every candidate that reached N — is re-measured exactly in mpf, and
// the reference's rule (first longest in scan order) is applied to those exact lengths.
// The winner is the exhaustive search's whenever that one is on the short-list (every fixture tested);
// probe_search = 0 selects the exhaustive search.
// `spec` (optional): the caller's next step is the reference build at the winner. The winner by the GPU's counts is,
// more often than not, the winner by the exact lengths as well, so its build starts on a thread of its own next to the
// exact check; if the check confirms that candidate, *spec holds its tables and *spec_valid is set (same function, same
// inputs: the tables are what the caller would have built), else the speculative build is discarded.
void find_probe_assisted(newman_b200::Engine& eng, const ViewHP& v, const RoundParams& rp, int threads, int cmode,
                         const std::vector<uint8_t>& mask, int& row, int& col, int& length, int* n_exact,
                         newman_b200::FrameInfo& info, DeepTablesHost* spec = nullptr, bool* spec_valid = nullptr) {
  if (spec_valid) *spec_valid = false;
  std::vector<std::pair<int, int> > cand;
  newman_b200::probe_candidates(v, cand);
  const int n = (int)cand.size();
  std::vector<int32_t> pix((size_t)n);
  for (int i = 0; i < n; i++) pix[i] = cand[i].first * v.nc + cand[i].second;
  std::vector<int32_t> uniq(pix);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());

  HostTrace tr("findProbe");
  DeepTablesHost T;
  newman_b200::build_tables(v, v.nr / 2, v.nc / 2, T, threads);
  tr.lap("centre reference");
  newman_b200::FrameInfo scratch;
  // orbit lengths are wanted, not the user's speed/accuracy trade-off: never a looser series tolerance
  // than the reference's default (mandelbrot.cpp:9)
  RoundParams rq = rp;
  if (!(rq.tol <= 1e-10)) rq.tol = 1e-10;
  // glitched candidates are finished by the rebasing pass, not against a secondary reference: the counts only rank the
  // candidates (the short-list is re-measured exactly below), and a reference build is most of a deep frame's host time
  rq.max_secondary = 0;
  run_rounds(eng, v, rq, T, cmode, mask, &uniq, scratch);  // cardioid/bulb candidates: (N, 0) without iterating
  info.probe_iters += scratch.executed_iters;  // kept apart from the frame's own counters
  std::vector<nm_escape> got((size_t)n);
  eng.check(nm_read_pixels(eng.ctx, pix.data(), n, got.data()), "nm_read_pixels");
  tr.lap("candidates on the GPU");

  const int kShort = 16;
  int maxc = 0;
  for (int i = 0; i < n; i++) if (got[i].iterations > maxc) maxc = got[i].iterations;
  std::vector<int> order((size_t)n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return got[a].iterations > got[b].iterations; });
  const int floor_count = maxc - (int)(0.002 * maxc) - 2;
  std::vector<int> which;
  for (int k = 0; k < n; k++) {
    const int i = order[k];
    if (k < kShort || got[i].iterations >= floor_count || got[i].iterations >= rp.N) which.push_back(i);
    else break;
  }
  std::sort(which.begin(), which.end());  // scan order
  // A candidate that never escapes cannot be beaten by a later one (strict '>'): of those the GPU
  // reports at N only the first in scan order needs the exact check — if it confirms.
  int first_full = -1;
  for (int i : which) if (got[i].iterations >= rp.N) { first_full = i; break; }
  if (first_full >= 0) {
    std::vector<int> one(1, first_full), l1;
    newman_b200::probe_lengths(v, cand, one, 1, l1);
    if (l1[0] < rp.N) {  // finite-precision artefact of the exact orbit: let the exhaustive search decide
      newman_b200::find_probe(v, threads, row, col, length);
      if (n_exact) *n_exact = n;
      return;
    }
    std::vector<int> keep;
    for (int i : which) if (i <= first_full && (got[i].iterations < rp.N || i == first_full)) keep.push_back(i);
    which.swap(keep);
  }
  struct SpecBuild {      // joined on every way out of this function
    std::thread th;
    DeepTablesHost T;
    std::string fail;
    int idx = -1;
    ~SpecBuild() { if (th.joinable()) th.join(); }
  } sb;
  // A wrong guess costs the exact check some of its cores and the wait for the discarded build (measured: cfg3, whose
  // ranking is ill-conditioned, +32 ms per frame; cfg2, where the guess holds, -6 ms): after a miss the next few frames
  // of the process do not speculate.
  static std::atomic<int> skip_after_miss(0);
  bool speculate = spec && spec_valid && !which.empty() && getenv("NM_NO_SPECULATION") == nullptr;
  if (speculate && skip_after_miss.load(std::memory_order_relaxed) > 0) {
    skip_after_miss.fetch_sub(1, std::memory_order_relaxed);
    speculate = false;
  }
  if (speculate) {
    size_t b = 0;   // first longest in scan order, by the GPU's counts
    for (size_t k = 1; k < which.size(); k++)
      if (got[which[k]].iterations > got[which[b]].iterations) b = k;
    sb.idx = which[b];
    const int sr = cand[sb.idx].first, sc = cand[sb.idx].second;
    SpecBuild* psb = &sb;
    const ViewHP* pv = &v;
    sb.th = std::thread([psb, pv, sr, sc, threads]() {
      try { newman_b200::build_tables(*pv, sr, sc, psb->T, threads); }
      catch (const std::exception& e) { psb->fail = e.what(); }
      catch (...) { psb->fail = "speculative reference build failed"; }
    });
  }
  std::vector<int> len;
  newman_b200::probe_lengths(v, cand, which, threads, len);
  // Consistency guard. The short-list is only as good as the perturbation counts that ranked it: if the exact (mpf)
  // lengths differ from their candidates' counts by more than the list's own width, a candidate outside the list could
  // be the exhaustive search's winner. Widen the list once — everything within twice the observed disagreement of the
  // maximum, at most kWide candidates by rank — and report whether that covered it (FrameInfo::probe_consistent = 0:
  // the criterion is ill-conditioned on this view — e.g. 1e-100, where the exhaustive winner's mpf orbit is 6 % longer
  // than its perturbation count — and only probe_search = 0 is guaranteed to return the reference's probe).
  info.probe_consistent = 1;
  if (first_full < 0) {
    const int kWide = 256;
    const int margin = maxc - floor_count;
    // (the MEDIAN disagreement: single candidates whose last iterations are chaotic differ by hundreds of iterations on
    // any deep view — DESIGN.md section 6 — and say nothing about the ranking as a whole)
    std::vector<int> devs;
    for (size_t k = 0; k < which.size(); k++)
      devs.push_back(len[k] > got[which[k]].iterations ? len[k] - got[which[k]].iterations : got[which[k]].iterations - len[k]);
    std::sort(devs.begin(), devs.end());
    const int max_dev = devs.empty() ? 0 : devs[devs.size() / 2];
    if (max_dev > margin) {
      const int floor2 = maxc - 2 * max_dev - margin;
      std::vector<char> have((size_t)n, 0);
      for (int i : which) have[(size_t)i] = 1;
      std::vector<int> more;
      bool covered = true;
      for (int k = 0; k < n; k++) {
        const int i = order[k];
        if (got[i].iterations < floor2) break;
        if (have[(size_t)i]) continue;
        if ((int)(which.size() + more.size()) >= kWide) { covered = false; break; }
        more.push_back(i);
      }
      // a list that cannot be made wide enough is not widened at all (each exact orbit costs as much as a third of a
      // reference build): the frame keeps the short-list's winner and says so
      if (!covered) more.clear();
      if (!more.empty()) {
        std::vector<int> len2;
        newman_b200::probe_lengths(v, cand, more, threads, len2);
        std::vector<std::pair<int, int> > all;   // (candidate, exact length), then back into scan order
        for (size_t k = 0; k < which.size(); k++) all.push_back(std::make_pair(which[k], len[k]));
        for (size_t k = 0; k < more.size(); k++) all.push_back(std::make_pair(more[k], len2[k]));
        std::sort(all.begin(), all.end());
        which.clear(); len.clear();
        for (const std::pair<int, int>& p : all) { which.push_back(p.first); len.push_back(p.second); }
      }
      info.probe_consistent = covered ? 1 : 0;
    }
  }
  tr.lap("exact check");
  if (n_exact) *n_exact = (int)which.size();
  info.probe_exact = (unsigned long long)which.size();
  size_t best = 0;
  for (size_t k = 1; k < which.size(); k++)
    if (len[k] > len[best]) best = k;  // first longest wins (strict '>' at mandelbrot.cpp:90)
  row = cand[which[best]].first;
  col = cand[which[best]].second;
  length = len[best];
  if (sb.th.joinable()) {
    sb.th.join();
    if (sb.idx == which[best] && sb.fail.empty()) {
      *spec = std::move(sb.T);
      *spec_valid = true;
    }
    if (!*spec_valid) skip_after_miss.store(8, std::memory_order_relaxed);
    tr.lap(*spec_valid ? "speculative reference (kept)" : "speculative reference (discarded)");
  }
}
// Exact mode (Mandelbrot::exact). After the frame: (1) keep its raster, (2) render the frame once more against the orbit
// TRUNCATED to doubles instead of rounded to nearest — same probe, same rounds —, (3) the samples whose escape COUNT differs
// between the two are the ones whose last iterations amplify FP64's 5e-15 past 1, (4) the first raster
// comes back and those samples are repeated from K2's hand-over in double-double arithmetic against the primary reference
// (k3_dd.cuh). On cfg2's 6 144-sample adjudication set the probe flags 72 samples, among them every one of the 26 the FP64
// frame gets wrong, and the double-double pass reproduces the reference's count on all of them.
void refine_exact(newman_b200::Engine& eng, const ViewHP& v, const RoundParams& rp, const DeepTablesHost& T0, int cmode,
                  const std::vector<uint8_t>& mask, const std::vector<int32_t>& primary_glitched, newman_b200::FrameInfo& info) {
  nm_ctx* ctx = eng.ctx;
  eng.check(nm_raster_keep(ctx), "nm_raster_keep");
  newman_b200::FrameInfo probe;
  {
    DeepTablesHost Tb = T0;
    RoundParams rq = rp;
    rq.orbit_trunc = true;
    run_rounds(eng, v, rq, Tb, cmode, mask, nullptr, probe);
  }
  info.refine_ms += probe.device_ms;
  info.kernel_launches += probe.kernel_launches;
  info.host_precompute_s += probe.host_precompute_s;
  int64_t n = nm_raster_diff(ctx, nullptr, 0);
  if (n < 0) eng.check((int)n, "nm_raster_diff");
  std::vector<int32_t> list((size_t)n);
  if (n > 0) {
    const int64_t m = nm_raster_diff(ctx, list.data(), n);
    if (m != n) eng.check(m < 0 ? (int)m : NM_ESTATE, "nm_raster_diff");
  }
  eng.check(nm_raster_restore(ctx), "nm_raster_restore");
  // ... plus every sample the primary round flagged as glitched: the frame finished those against a secondary reference
  // or by rebasing, i.e. from ANOTHER series start than the reference's own algorithm takes (its series is relative to its
  // probe = our primary reference), which both renderings do alike — the probe cannot see it. In double-double the
  // cancellation behind the glitch flag is harmless, so they are repeated against the primary reference like the rest.
  list.insert(list.end(), primary_glitched.begin(), primary_glitched.end());
  std::sort(list.begin(), list.end());
  list.erase(std::unique(list.begin(), list.end()), list.end());
  n = (int64_t)list.size();
  info.refined = (unsigned long long)n;
  if (n == 0) return;
  const int fe = T0.finite ? (rp.force_floatexp == 1 ? 1 : 0) : 1;
  nm_deep_tables t;
  memset(&t, 0, sizeof t);
  t.M = T0.M; t.N = rp.N; t.has_escape = T0.has_escape ? 1 : 0;
  t.tol = rp.tol; t.glitch_tol = rp.gtol;
  t.x_hi = T0.x_hi.data(); t.x_lo = T0.x_lo.data();
  t.a = fe ? T0.a_m.data() : T0.a.data(); t.b = fe ? T0.b_m.data() : T0.b.data(); t.c = fe ? T0.c_m.data() : T0.c.data();
  t.a_exp = fe ? T0.a_e.data() : nullptr; t.b_exp = fe ? T0.b_e.data() : nullptr; t.c_exp = fe ? T0.c_e.data() : nullptr;
  t.eps_re_lo = T0.eps_re_lo.data(); t.eps_im_lo = T0.eps_im_lo.data();
  eng.check(nm_frame_deep(ctx, &t, T0.eps_re.data(), v.nc, T0.eps_im.data(), v.nr, NM_CARDIOID_NONE, nullptr, list.data(), n,
                          NM_MODE_DD),
            "nm_frame_deep (double-double)");
  eng.check(nm_launch(ctx), "nm_launch");
  nm_stats st; eng.check(nm_frame_stats(ctx, &st), "nm_frame_stats");
  info.refine_ms += st.ms_k2 + st.ms_k3;
  info.refine_iters += st.executed_iters;
  info.kernel_launches += st.kernel_launches;
  info.fixups += st.fixups;
}
}  // namespace


// ---- one frame over several GPUs ---------------------------------------------------------------------------------
// Every rank of a render group (multi_host.h) runs this with a Mandelbrot in the same state; nothing of `m` is written.
// Rank r renders the blocks of `band` grid rows b = r, r + world, ... of the raster. Rank 0 does the host
// arbitrary-precision work (cardioid classification, probe search on its own GPU, orbit + series of every reference) and
// broadcasts each reference's tables; the choice of the next secondary reference is the single-GPU rule (earliest flag
// iteration, lowest sample id on ties) taken over all ranks by one 8-byte MIN all-reduce, so the N-GPU raster is
// byte-identical to the 1-GPU one. out / mode: see nmm_render (include/newman_b200.h).
namespace newman_b200 {

namespace {
struct FrameHeader {   // rank 0 -> all, once per reference
  int32_t kind;        // 2: a deep reference follows; 0: rank 0 failed (text in `msg`)
  int32_t M, has_escape, fe, cmode, mask_bytes_follow;
  uint64_t blob_bytes;
  char msg[160];
};
size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
}  // namespace

void render_collective(RankLink& link, Mandelbrot& m, int band, nm_escape* out, int mode, FrameInfo& info,
                       const std::atomic<bool>* cancel) {
  const double t_begin = now_s();
  nm_ctx* ctx = link.ctx;
  info = FrameInfo();
  ViewHP v;
  v.center_re = m.center.re.get_mpf_t(); v.center_im = m.center.im.get_mpf_t();
  v.sz_re = m.sz.re.get_mpf_t(); v.sz_im = m.sz.im.get_mpf_t();
  v.nr = m.rows(); v.nc = m.cols(); v.N = m.N;
  v.prec = max_prec(m.center, m.sz);
  info.precision_bits = (int)v.prec;
  if (band < 1 || v.nr % band) throw std::runtime_error("newman_b200: band_rows must divide the number of grid rows");
  const int n_blocks = v.nr / band;
  const int nb_loc = RankLink::blocks_of(link.rank, link.world, n_blocks);
  const int nr_loc = nb_loc * band;
  auto global_row = [&](int r_loc) { return ((r_loc / band) * link.world + link.rank) * band + r_loc % band; };
  auto check = [&](int rc, const char* what) {
    if (rc != NM_OK) throw std::runtime_error(std::string("newman_b200: ") + what + ": " + nm_last_error(ctx));
  };
  auto absorb = [&](const nm_stats& st) {
    info.executed_iters += st.executed_iters; info.series_evals += st.series_evals;
    info.skipped_pixels += st.skipped_pixels; info.rebased += st.rebased; info.fixups += st.fixups;
    info.kernel_launches += st.kernel_launches;
    info.device_ms += st.ms_k1 + st.ms_k2 + st.ms_k3;
  };

  if (m.useHardware()) {
    // plain double: the coordinates are a few thousand mpf operations — every rank forms its own, nothing is exchanged
    info.hardware = true;
    std::vector<double> c_re, c_im, c_im_loc((size_t)(nr_loc > 0 ? nr_loc : 1));
    pixel_coords(v, c_re, c_im);
    for (int r = 0; r < nr_loc; r++) c_im_loc[(size_t)r] = c_im[(size_t)global_row(r)];
    info.host_precompute_s = now_s() - t_begin;
    if (nr_loc > 0) {
      check(nm_frame_hw(ctx, c_re.data(), v.nc, c_im_loc.data(), nr_loc, m.N), "nm_frame_hw");
      check(nm_launch(ctx), "nm_launch");
      int64_t n_amb = nm_frame_ambiguous(ctx, nullptr, 0);
      if (n_amb < 0) check((int)n_amb, "nm_frame_ambiguous");
      if (n_amb > 0) {
        std::vector<int32_t> amb((size_t)n_amb);
        nm_frame_ambiguous(ctx, amb.data(), n_amb);
        for (int32_t pix : amb)
          if (in_cardioid_pixel(v, global_row(pix / v.nc), pix % v.nc)) {
            nm_escape e; e.iterations = m.N; e.smoothing = 0.0f;
            check(nm_poke(ctx, pix, e), "nm_poke");
          }
        info.ambiguous = (unsigned long long)n_amb;
      }
      nm_stats st; check(nm_frame_stats(ctx, &st), "nm_frame_stats"); absorb(st);
    }
  } else {
    RoundParams rp = {m.N, m.max_secondary, m.force_floatexp, m.error_tolerance, m.glitch_tolerance, m.host_threads};
    DeepTablesHost T;
    std::vector<uint8_t> mask;
    int cmode = NM_CARDIOID_NONE;
    std::string fail;
    if (link.rank == 0) {
      try {
        cmode = classify_cardioid(v, m.host_threads, mask);
        if (cmode == NM_CARDIOID_ALL) {
          ViewHP v1 = v; v1.N = 1;
          build_tables(v1, v.nr / 2, v.nc / 2, T, 1);
        } else {
          int prow, pcol, plen;
          if (m.probe_search == 0) find_probe(v, m.host_threads, prow, pcol, plen);
          else {
            Engine probe_eng(ctx);   // borrows the rank's context for the candidate frame
            find_probe_assisted(probe_eng, v, rp, m.host_threads, cmode, mask, prow, pcol, plen, nullptr, info);
          }
          build_tables(v, prow, pcol, T, m.host_threads);
        }
        info.orbit_len = T.M; info.probe_row = T.probe_row; info.probe_col = T.probe_col;
      } catch (const std::exception& e) { fail = e.what(); }
      info.host_precompute_s = now_s() - t_begin;
    }
    std::vector<int32_t> rq_pix, rq_iter;
    for (int round = 0;; round++) {
      // ---- rank 0 packs the reference (header + one blob), everyone receives it --------------------------------
      FrameHeader h;
      memset(&h, 0, sizeof h);
      size_t off_xhi = 0, off_xlo = 0, off_a = 0, off_b = 0, off_c = 0, off_ae = 0, off_be = 0, off_ce = 0, off_er = 0, off_ei = 0,
             off_ere = 0, off_eie = 0, off_mask = 0;
      auto layout = [&](int M, int he, int fe, bool with_mask) {
        size_t o = 0;
        auto take = [&](size_t bytes) { const size_t at = o; o = align16(o + bytes); return at; };
        off_xhi = take((size_t)2 * (M + he) * 8); off_xlo = take((size_t)2 * M * 8);
        off_a = take((size_t)2 * M * 8); off_b = take((size_t)2 * M * 8); off_c = take((size_t)2 * M * 8);
        if (fe) { off_ae = take((size_t)2 * M * 4); off_be = take((size_t)2 * M * 4); off_ce = take((size_t)2 * M * 4); }
        off_er = take((size_t)v.nc * 8); off_ei = take((size_t)v.nr * 8);
        if (fe == 2) { off_ere = take((size_t)v.nc * 4); off_eie = take((size_t)v.nr * 4); }
        if (with_mask) off_mask = take((size_t)v.nr * v.nc);
        return o;
      };
      const char* blob_host = nullptr;
      if (link.rank == 0) {
        if (!fail.empty()) { h.kind = 0; snprintf(h.msg, sizeof h.msg, "%s", fail.c_str()); }
        else {
          int fe = T.finite ? 0 : 1;
          if (T.pitch_exp < -380) fe = 2;
          if (rp.force_floatexp > fe) fe = rp.force_floatexp > 2 ? 2 : rp.force_floatexp;
          h.kind = 2; h.M = T.M; h.has_escape = T.has_escape ? 1 : 0; h.fe = fe; h.cmode = round == 0 ? cmode : NM_CARDIOID_NONE;
          h.mask_bytes_follow = (round == 0 && cmode == NM_CARDIOID_MASK) ? 1 : 0;
          h.blob_bytes = layout(h.M, h.has_escape, fe, h.mask_bytes_follow != 0);
          char* b = (char*)link.pinned(h.blob_bytes);
          memcpy(b + off_xhi, T.x_hi.data(), (size_t)2 * (h.M + h.has_escape) * 8);
          memcpy(b + off_xlo, T.x_lo.data(), (size_t)2 * h.M * 8);
          memcpy(b + off_a, fe ? T.a_m.data() : T.a.data(), (size_t)2 * h.M * 8);
          memcpy(b + off_b, fe ? T.b_m.data() : T.b.data(), (size_t)2 * h.M * 8);
          memcpy(b + off_c, fe ? T.c_m.data() : T.c.data(), (size_t)2 * h.M * 8);
          if (fe) {
            memcpy(b + off_ae, T.a_e.data(), (size_t)2 * h.M * 4); memcpy(b + off_be, T.b_e.data(), (size_t)2 * h.M * 4);
            memcpy(b + off_ce, T.c_e.data(), (size_t)2 * h.M * 4);
          }
          memcpy(b + off_er, fe == 2 ? T.eps_re_m.data() : T.eps_re.data(), (size_t)v.nc * 8);
          memcpy(b + off_ei, fe == 2 ? T.eps_im_m.data() : T.eps_im.data(), (size_t)v.nr * 8);
          if (fe == 2) { memcpy(b + off_ere, T.eps_re_e.data(), (size_t)v.nc * 4); memcpy(b + off_eie, T.eps_im_e.data(), (size_t)v.nr * 4); }
          if (h.mask_bytes_follow) memcpy(b + off_mask, mask.data(), (size_t)v.nr * v.nc);
          blob_host = b;
        }
      }
      link.bcast_host(&h, sizeof h);
      if (h.kind == 0) throw std::runtime_error(std::string("newman_b200: rank 0 failed: ") + h.msg);
      layout(h.M, h.has_escape, h.fe, h.mask_bytes_follow != 0);
      const char* d = (const char*)link.bcast_device(blob_host, h.blob_bytes, 0);
      info.floatexp = h.fe;
      if (round == 0) { info.orbit_len = h.M; cmode = h.cmode; }

      // ---- this rank's bands against the reference -------------------------------------------------------------
      const bool last = round >= rp.max_secondary;
      const bool listed = round > 0;
      if (nr_loc > 0 && (!listed || !rq_pix.empty())) {
        nm_deep_tables t;
        memset(&t, 0, sizeof t);
        t.M = h.M; t.N = rp.N; t.has_escape = h.has_escape; t.tol = rp.tol; t.glitch_tol = rp.gtol;
        t.x_hi = (const double*)(d + off_xhi); t.x_lo = (const double*)(d + off_xlo);
        t.a = (const double*)(d + off_a); t.b = (const double*)(d + off_b); t.c = (const double*)(d + off_c);
        if (h.fe) { t.a_exp = (const int32_t*)(d + off_ae); t.b_exp = (const int32_t*)(d + off_be); t.c_exp = (const int32_t*)(d + off_ce); }
        const double* eps_im_loc = (const double*)link.gather_rows(d + off_ei, 8, band, n_blocks, 0);
        if (h.fe == 2) {
          t.eps_re_exp = (const int32_t*)(d + off_ere);
          t.eps_im_exp = (const int32_t*)link.gather_rows(d + off_eie, 4, band, n_blocks, 1);
        }
        const uint8_t* mask_loc = nullptr;
        if (h.mask_bytes_follow) mask_loc = (const uint8_t*)link.gather_rows(d + off_mask, (size_t)v.nc, band, n_blocks, 2);
        check(nm_frame_deep(ctx, &t, (const double*)(d + off_er), v.nc, eps_im_loc, nr_loc, listed ? NM_CARDIOID_NONE : h.cmode,
                            mask_loc, listed ? rq_pix.data() : nullptr, listed ? (int64_t)rq_pix.size() : 0,
                            last ? NM_MODE_REBASE : NM_MODE_REQUEUE),
              "nm_frame_deep");
        check(nm_launch(ctx), "nm_launch");
        nm_stats st; check(nm_frame_stats(ctx, &st), "nm_frame_stats"); absorb(st);
        int64_t n_rq = nm_frame_requeue(ctx, nullptr, nullptr, 0);
        if (n_rq < 0) check((int)n_rq, "nm_frame_requeue");
        rq_pix.resize((size_t)n_rq); rq_iter.resize((size_t)n_rq);
        if (n_rq) nm_frame_requeue(ctx, rq_pix.data(), rq_iter.data(), n_rq);
        info.glitched += (unsigned long long)n_rq;
      } else {
        rq_pix.clear(); rq_iter.clear();
      }
      info.references++;
      // ---- next reference: the glitched sample flagged earliest, lowest (global) sample id on ties, over all ranks ----
      uint64_t key = ~(uint64_t)0;
      for (size_t i = 0; i < rq_pix.size(); i++) {
        const uint64_t g = (uint64_t)global_row(rq_pix[i] / v.nc) * (uint64_t)v.nc + (uint64_t)(rq_pix[i] % v.nc);
        const uint64_t k = ((uint64_t)rq_iter[i] << 40) | g;
        if (k < key) key = k;
      }
      // a cancel request (Mandelbrot::cancel, any thread) travels with the same reduction, so every rank leaves the frame
      // at the same point and nobody is left waiting in a collective
      if (cancel && cancel->load()) key = 0;
      key = link.allreduce_min(key);
      if (key == 0) { info.cancelled = true; break; }
      if (key == ~(uint64_t)0) break;
      if (link.rank == 0) {
        const uint64_t g = key & (((uint64_t)1 << 40) - 1);
        const double t_hp = now_s();
        try { build_tables(v, (int)(g / (uint64_t)v.nc), (int)(g % (uint64_t)v.nc), T, rp.host_threads); }
        catch (const std::exception& e) { fail = e.what(); }
        info.host_precompute_s += now_s() - t_hp;
      }
    }
  }

  // ---- bands back to the host raster ---------------------------------------------------------------------------
  if (info.cancelled) {
    info.frame_s = now_s() - t_begin;
    return;
  }
  if (out || mode == NMM_RETURN_LOCAL) {
    const size_t block_bytes = (size_t)band * v.nc * sizeof(nm_escape);
    void* bd = link.band_buffer((size_t)(nr_loc > 0 ? nr_loc : 1) * v.nc * sizeof(nm_escape));
    if (nr_loc > 0) check(nm_read_rows(ctx, 0, nr_loc, (nm_escape*)bd), "nm_read_rows");
    link.return_band(bd, block_bytes, n_blocks, out, mode);
  } else if (link.world > 1 && mode == NMM_RETURN_ROOT) {
    const size_t block_bytes = (size_t)band * v.nc * sizeof(nm_escape);
    void* bd = link.band_buffer((size_t)(nr_loc > 0 ? nr_loc : 1) * v.nc * sizeof(nm_escape));
    if (nr_loc > 0) check(nm_read_rows(ctx, 0, nr_loc, (nm_escape*)bd), "nm_read_rows");
    link.return_band(bd, block_bytes, n_blocks, nullptr, mode);
  }
  // ---- counters over the group ----------------------------------------------------------------------------------
  uint64_t sums[8] = {info.executed_iters, info.series_evals, info.skipped_pixels, info.glitched, info.rebased, info.fixups,
                      info.kernel_launches, info.ambiguous};
  link.allreduce_sum(sums, 8);
  info.executed_iters = sums[0]; info.series_evals = sums[1]; info.skipped_pixels = sums[2]; info.glitched = sums[3];
  info.rebased = sums[4]; info.fixups = sums[5]; info.kernel_launches = sums[6]; info.ambiguous = sums[7];
  info.device_ms = (double)link.allreduce_max((uint64_t)(info.device_ms * 1e3)) * 1e-3;
  info.frame_s = now_s() - t_begin;
}

}  // namespace newman_b200


namespace newman_b200 {
// The render group behind `Mandelbrot::devices`: one RankLink (context + NCCL rank) per GPU, created by one thread per
// GPU (ncclCommInitRank is a rendezvous), and one thread per GPU again for every frame.
class Group {
public:
  std::vector<int> devices;
  std::vector<std::unique_ptr<RankLink> > links;
  std::atomic<bool> cancel_requested{false};
  explicit Group(const std::vector<int>& devs) : devices(devs), links(devs.size()) {
    uint8_t id[NMM_ID_BYTES];
    RankLink::unique_id(id);
    std::vector<std::string> errs(devs.size());
    std::vector<std::thread> th;
    for (size_t r = 0; r < devs.size(); r++)
      th.emplace_back([&, r]() {
        try { links[r].reset(new RankLink(devs[r], (int)r, (int)devs.size(), id)); }
        catch (const std::exception& e) { errs[r] = e.what(); }
      });
    for (std::thread& t : th) t.join();
    for (const std::string& e : errs) if (!e.empty()) throw std::runtime_error(e);
  }
  void render(Mandelbrot& m, int band, nm_escape* out, FrameInfo& info) {
    cancel_requested.store(false);
    std::vector<std::string> errs(links.size());
    std::vector<FrameInfo> infos(links.size());
    std::vector<std::thread> th;
    for (size_t r = 0; r < links.size(); r++)
      th.emplace_back([&, r]() {
        try { render_collective(*links[r], m, band, out, NMM_RETURN_LOCAL, infos[r], &cancel_requested); }
        catch (const std::exception& e) { errs[r] = e.what(); }
      });
    for (std::thread& t : th) t.join();
    for (const std::string& e : errs) if (!e.empty()) throw std::runtime_error(e);
    info = infos[0];
  }
};
}  // namespace newman_b200

Mandelbrot::Mandelbrot() : Mandelbrot(1, 1) {}

Mandelbrot::Mandelbrot(int nr, int nc)
    : grid(nr, nc), error_tolerance(1e-10), N(256), glitch_tolerance(1e-6), max_secondary(1), device(0), band_rows(4), host_threads(0),
      probe_search(1), exact(0), force_floatexp(0) {
  // default full view (mandelbrot.cpp:13-14)
  center.re = -0.5;
  center.im = 0.0;
  sz.re = 4.0 / nc;
  sz.im = 3.0 / nr;
  setPrecision();
}

void Mandelbrot::setPrecision() {
  const int bits = newman_b200::precision_bits_for(sz.re.get_mpf_t());
  mpf_set_default_prec(bits);  // global, like the reference (mandelbrot.cpp:47): callers' temporaries follow it
  center.re.set_prec(bits);
  center.im.set_prec(bits);
  sz.re.set_prec(bits);
  sz.im.set_prec(bits);
  rendered_.reset();  // the reference clears X/A/B/C here; our equivalent is dropping the frame
}

void Mandelbrot::loadLegacy(const char* fn) {
  std::ifstream in(fn);
  if (!in) return;  // the reference ignores a missing file too (mandelbrot.cpp:20-21)
  std::string cre, cim, sre, sim;
  int n = N;
  in >> n >> cre >> cim >> sre >> sim;
  if (!in) return;
  N = n;
  // parsed at the CURRENT precision of each field (mandelbrot.cpp:24-27), then rescaled from the
  // 800x600 window the file format assumes (30-31)
  center.re = cre.c_str();
  center.im = cim.c_str();
  sz.re = sre.c_str();
  sz.im = sim.c_str();
  sz.re = sz.re * (800.0 / cols());
  sz.im = sz.im * (600.0 / rows());
  setPrecision();
}

void Mandelbrot::load(const char* fn) { loadLegacy(fn); }

void Mandelbrot::save(const char* fn) {
  // the 5-line text format FractalViewer::save writes (viewer.cpp:12-23): N, centre, sz as they are
  // (loadLegacy's 800/cols rescale is the identity for the viewer's 800x600 window)
  FILE* fp = fopen(fn, "w");
  if (!fp) return;
  fprintf(fp, "%d\n", N);
  const mpf_class* fields[4] = {&center.re, &center.im, &sz.re, &sz.im};
  for (const mpf_class* f : fields) {
    mpf_out_str(fp, 10, 0, f->get_mpf_t());
    fprintf(fp, "\n");
  }
  fclose(fp);
}

bool Mandelbrot::useHardware() {
  const double minpreview = 1.5e-16;  // mandelbrot.cpp:257
  return sz.re.get_d() >= minpreview && sz.im.get_d() >= minpreview;
}

bool Mandelbrot::frameCurrent() const {
  if (!rendered_) return false;
  const Signature& s = *rendered_;
  return s.N == N && s.nr == grid.nr && s.nc == grid.nc && s.max_secondary == max_secondary &&
         s.tol == error_tolerance && s.gtol == glitch_tolerance && s.cre == center.re && s.cim == center.im &&
         s.sre == sz.re && s.sim == sz.im;
}

void Mandelbrot::precompute() { renderFrame(); }

void Mandelbrot::computeRow(int r) {
  (void)r;
  if (!frameCurrent()) renderFrame();  // rows of a current frame are already in `grid`
}

void Mandelbrot::cancel() {
  // any thread: the frame in flight is abandoned — persistent CTAs poll the flag (nm_cancel), the host loop checks it
  // between stages — and precompute() returns with frameInfo().cancelled set; the raster is then not current
  std::shared_ptr<newman_b200::Engine> e = engine_;
  if (e) { e->cancel_requested.store(true); nm_cancel(e->ctx); }
  std::shared_ptr<newman_b200::Group> g = group_;
  if (g) g->cancel_requested.store(true);
}

void Mandelbrot::renderFrame() {
  try {
    renderFrameImpl();
  } catch (const newman_b200::Cancelled&) {
    info_.cancelled = true;
    rendered_.reset();
  }
  if (engine_) engine_->cancel_requested.store(false);   // a request only ever concerns the frame in flight
}

void Mandelbrot::renderFrameImpl() {
  const double t_begin = now_s();
  if (!devices.empty()) {
    // one frame over several GPUs (render_collective above): a thread, a context and an NCCL rank per GPU
    if (!group_ || group_->devices != devices) group_ = std::make_shared<newman_b200::Group>(devices);
    int band = band_rows > 0 ? band_rows : 1;
    while (grid.nr % band) band--;   // any raster renders; bands shrink to the largest divisor of the row count
    group_->render(*this, band, reinterpret_cast<nm_escape*>(grid.values.data()), info_);
    if (info_.cancelled) { rendered_.reset(); info_.frame_s = now_s() - t_begin; return; }
    std::shared_ptr<Signature> sg = std::make_shared<Signature>();
    sg->N = N; sg->nr = grid.nr; sg->nc = grid.nc; sg->max_secondary = max_secondary;
    sg->tol = error_tolerance; sg->gtol = glitch_tolerance;
    sg->cre = mpf_class(center.re, center.re.get_prec()); sg->cim = mpf_class(center.im, center.im.get_prec());
    sg->sre = mpf_class(sz.re, sz.re.get_prec()); sg->sim = mpf_class(sz.im, sz.im.get_prec());
    rendered_ = sg;
    info_.frame_s = now_s() - t_begin;
    return;
  }
  if (!engine_ || engine_->device != device) engine_ = std::make_shared<newman_b200::Engine>(device);
  newman_b200::Engine& eng = *engine_;
  nm_ctx* ctx = eng.ctx;
  info_ = newman_b200::FrameInfo();
  eng.cancel_requested.store(false);

  ViewHP v;
  v.center_re = center.re.get_mpf_t(); v.center_im = center.im.get_mpf_t();
  v.sz_re = sz.re.get_mpf_t(); v.sz_im = sz.im.get_mpf_t();
  v.nr = grid.nr; v.nc = grid.nc; v.N = N;
  v.prec = max_prec(center, sz);
  info_.precision_bits = (int)v.prec;
  nm_escape* out = reinterpret_cast<nm_escape*>(grid.values.data());

  auto absorb = [&](const nm_stats& st) {
    info_.executed_iters += st.executed_iters; info_.series_evals += st.series_evals;
    info_.skipped_pixels += st.skipped_pixels; info_.rebased += st.rebased; info_.fixups += st.fixups;
    info_.kernel_launches += st.kernel_launches;
    info_.device_ms += st.ms_k1 + st.ms_k2 + st.ms_k3;
  };

  if (useHardware()) {
    info_.hardware = true;
    std::vector<double> c_re, c_im;
    newman_b200::pixel_coords(v, c_re, c_im);
    info_.host_precompute_s = now_s() - t_begin;
    eng.check(nm_frame_hw(ctx, c_re.data(), v.nc, c_im.data(), v.nr, N), "nm_frame_hw");
    eng.check(nm_launch(ctx), "nm_launch");
    int64_t n_amb = nm_frame_ambiguous(ctx, nullptr, 0);
    if (n_amb < 0) eng.check((int)n_amb, "nm_frame_ambiguous");
    if (n_amb > 0) {  // the reference's mpf test decides the samples double could not
      std::vector<int32_t> amb((size_t)n_amb);
      nm_frame_ambiguous(ctx, amb.data(), n_amb);
      for (int32_t pix : amb)
        if (newman_b200::in_cardioid_pixel(v, pix / v.nc, pix % v.nc)) {
          nm_escape e; e.iterations = N; e.smoothing = 0.0f;
          eng.check(nm_poke(ctx, pix, e), "nm_poke");
        }
      info_.ambiguous = (unsigned long long)n_amb;
    }
    eng.check(nm_read_rows(ctx, 0, v.nr, out), "nm_read_rows");
    nm_stats st; nm_frame_stats(ctx, &st); absorb(st);
    info_.references = 0;
  } else {
    HostTrace tr("renderFrame");
    std::vector<uint8_t> mask;
    int cmode = newman_b200::classify_cardioid(v, host_threads, mask);
    tr.lap("cardioid classification");
    DeepTablesHost T;
    if (cmode == NM_CARDIOID_ALL) {
      // every sample returns (N, 0) at mandelbrot.cpp:149-153: no probe search, a one-entry orbit
      ViewHP v1 = v;
      v1.N = 1;
      newman_b200::build_tables(v1, v.nr / 2, v.nc / 2, T, 1);
    } else {
      int prow, pcol, plen;
      RoundParams rp0 = {N, max_secondary, force_floatexp, error_tolerance, glitch_tolerance, host_threads};
      bool have = false;
      if (probe_search == 0) newman_b200::find_probe(v, host_threads, prow, pcol, plen);
      else find_probe_assisted(eng, v, rp0, host_threads, cmode, mask, prow, pcol, plen, nullptr, info_, &T, &have);
      tr.lap("findProbe (total)");
      if (!have) newman_b200::build_tables(v, prow, pcol, T, host_threads);
      tr.lap("primary reference");
    }
    info_.orbit_len = T.M; info_.probe_row = T.probe_row; info_.probe_col = T.probe_col;
    info_.host_precompute_s = now_s() - t_begin;
    info_.references = 0;
    RoundParams rp = {N, max_secondary, force_floatexp, error_tolerance, glitch_tolerance, host_threads};
    DeepTablesHost T0;
    const bool refine = exact != 0 && cmode != NM_CARDIOID_ALL;
    if (refine) T0 = T;   // run_rounds overwrites T with the secondary references
    std::vector<int32_t> primary_glitched;
    run_rounds(eng, v, rp, T, cmode, mask, nullptr, info_, &primary_glitched);
    tr.lap("rounds (GPU + secondary)");
    if (refine && info_.floatexp != 2) {
      refine_exact(eng, v, rp, T0, cmode, mask, primary_glitched, info_);
      tr.lap("exact mode (probe + dd)");
    }
    eng.check(nm_read_rows(ctx, 0, v.nr, out), "nm_read_rows");
    tr.lap("raster D2H");
  }

  std::shared_ptr<Signature> s = std::make_shared<Signature>();
  s->N = N; s->nr = grid.nr; s->nc = grid.nc; s->max_secondary = max_secondary;
  s->tol = error_tolerance; s->gtol = glitch_tolerance;
  s->cre = mpf_class(center.re, center.re.get_prec()); s->cim = mpf_class(center.im, center.im.get_prec());
  s->sre = mpf_class(sz.re, sz.re.get_prec()); s->sim = mpf_class(sz.im, sz.im.get_prec());
  rendered_ = s;
  info_.frame_s = now_s() - t_begin;
}

void Mandelbrot::findProbe(int& row, int& col, int& length, int* n_exact) {
  ViewHP v;
  v.center_re = center.re.get_mpf_t(); v.center_im = center.im.get_mpf_t();
  v.sz_re = sz.re.get_mpf_t(); v.sz_im = sz.im.get_mpf_t();
  v.nr = grid.nr; v.nc = grid.nc; v.N = N;
  v.prec = max_prec(center, sz);
  if (probe_search == 0) {
    newman_b200::find_probe(v, host_threads, row, col, length);
    if (n_exact) *n_exact = 3 * ((v.nc + 1) / 2) + (v.nr + 1) / 2;
    return;
  }
  if (!engine_ || engine_->device != device) engine_ = std::make_shared<newman_b200::Engine>(device);
  std::vector<uint8_t> mask;
  int cmode = newman_b200::classify_cardioid(v, host_threads, mask);
  RoundParams rp = {N, max_secondary, force_floatexp, error_tolerance, glitch_tolerance, host_threads};
  newman_b200::FrameInfo scratch;
  find_probe_assisted(*engine_, v, rp, host_threads, cmode, mask, row, col, length, n_exact, scratch);
  rendered_.reset();  // the device raster now holds the candidates, not a frame
}

void Mandelbrot::resolveRGB(const unsigned char* pal_rgb, int n_pal, int sc, bool smooth, unsigned char* rgb_out) {
  if (!engine_ || engine_->device != device) engine_ = std::make_shared<newman_b200::Engine>(device);
  // the host raster is authoritative (scaleUp/scaleDown and copies edit it), so resolve from it
  engine_->check(nm_resolve_grid(engine_->ctx, reinterpret_cast<const nm_escape*>(grid.values.data()), grid.nr, grid.nc,
                                 pal_rgb, n_pal, N, sc, smooth ? 1 : 0, rgb_out),
                 "nm_resolve_grid");
}

// ---- view geometry (mandelbrot.cpp:285-316): same mpf expressions, so views stay bit-identical ----
HPComplex Mandelbrot::pointAt(int r, int c, int sc) const {
  HPComplex pt;
  pt.re = center.re + (sc * c - cols() / 2) * sz.re;
  pt.im = center.im + (rows() / 2 - sc * r - 1) * sz.im;
  return pt;
}

void Mandelbrot::translate(int dr, int dc, int sc) {
  center.re = center.re - dc * sc * sz.re;
  center.im = center.im + dr * sc * sz.im;
}

void Mandelbrot::zoom(float scale) {
  const double inv = 1.0 / scale;
  sz.re *= inv;
  sz.im *= inv;
  setPrecision();
}

void Mandelbrot::zoomAt(float scale, int r, int c, int sc) {
  HPComplex anchor = pointAt(r, c, sc);  // the sample under the cursor keeps its coordinate
  const double inv = 1.0 / scale;
  sz.re *= inv;
  sz.im *= inv;
  center.re = anchor.re - (sc * c - cols() / 2) * sz.re;
  center.im = anchor.im - (rows() / 2 - sc * r - 1) * sz.im;
  setPrecision();
}

// ---- raster access and multisample rescale (mandelbrot.cpp:318-360) ---------------------------------
const RenderGrid::EscapeValue& Mandelbrot::at(int r, int c) { return grid.at(r, c); }

RenderGrid::EscapeValue Mandelbrot::at(int r, int c, int sc) {
  float sum = 0.0f;  // float32 accumulator, like the reference
  for (int r1 = 0; r1 < sc; r1++)
    for (int c1 = 0; c1 < sc; c1++) {
      const RenderGrid::EscapeValue& e = grid.at(sc * r + r1, sc * c + c1);
      sum += e.iterations + e.smoothing;
    }
  sum /= sc * sc;
  RenderGrid::EscapeValue avg;
  avg.iterations = (int)sum;
  avg.smoothing = sum - avg.iterations;
  return avg;
}

void Mandelbrot::scaleUp(int sc) {
  RenderGrid big(rows() * sc, cols() * sc);
  for (int r = 0; r < big.nr; r++)
    for (int c = 0; c < big.nc; c++) big.at(r, c) = grid.at(r / sc, c / sc);
  grid = std::move(big);
  sz.re = sz.re / sc;
  sz.im = sz.im / sc;
  setPrecision();
}

void Mandelbrot::scaleDown(int sc) {
  RenderGrid small_(rows() / sc, cols() / sc);
  for (int r = 0; r < small_.nr; r++)
    for (int c = 0; c < small_.nc; c++) small_.at(r, c) = at(r, c, sc);
  grid = std::move(small_);
  sz.re = sz.re * sc;
  sz.im = sz.im * sc;
  setPrecision();
}
