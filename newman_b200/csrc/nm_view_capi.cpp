// View-level C-ABI (nmv_*): one `class Mandelbrot` behind an opaque handle, so ctypes / cgo / JNI
// bindings drive the same drop-in object the reference's viewer would. No exception crosses.
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <exception>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/newman_b200.h"
#include "../../include/newman_b200/mandelbrot.h"
#include "hp_host.h"
#include "multi_host.h"

namespace {
struct Access : public Mandelbrot {  // reach the protected raster for bulk copies
  Access(int nr, int nc) : Mandelbrot(nr, nc) {}
  RenderGrid& g() { return grid; }
  const RenderGrid& g() const { return grid; }
};
}  // namespace

struct nmv_view {
  Access m;
  std::string err;
  newman_b200::DeepTablesHost tabs;
  nmv_view(int nr, int nc) : m(nr, nc) {}
};

struct nmm_rank {
  std::unique_ptr<newman_b200::RankLink> link;
  std::string err;
  int nr = 0, nc = 0, N = 0, band = 0;   // geometry of the last nmm_render (what nmm_resolve resolves)
};

namespace newman_b200 {
void render_collective(RankLink& link, Mandelbrot& m, int band, nm_escape* out, int mode, FrameInfo& info,
                       const std::atomic<bool>* cancel);   // mandelbrot_host.cpp
}

namespace {
std::string g_err;

void fill_info(const newman_b200::FrameInfo& f, nmv_frame_info* out) {
  out->hardware = f.hardware ? 1 : (f.floatexp ? 1 + f.floatexp : 0); out->precision_bits = f.precision_bits; out->orbit_len = f.orbit_len;
  out->probe_row = f.probe_row; out->probe_col = f.probe_col; out->references = f.references;
  out->executed_iters = f.executed_iters; out->series_evals = f.series_evals; out->skipped_pixels = f.skipped_pixels;
  out->probe_iters = f.probe_iters; out->probe_exact = f.probe_exact;
  out->glitched = f.glitched; out->rebased = f.rebased; out->fixups = f.fixups; out->kernel_launches = f.kernel_launches;
  out->ambiguous = f.ambiguous;
  out->host_precompute_s = f.host_precompute_s; out->device_ms = f.device_ms; out->frame_s = f.frame_s;
  out->probe_consistent = f.probe_consistent; out->cancelled = f.cancelled ? 1 : 0;
  out->refined = f.refined; out->refine_ms = f.refine_ms;
}

newman_b200::ViewHP hp_of(nmv_view* v) {
  newman_b200::ViewHP h;
  h.center_re = v->m.center.re.get_mpf_t(); h.center_im = v->m.center.im.get_mpf_t();
  h.sz_re = v->m.sz.re.get_mpf_t(); h.sz_im = v->m.sz.im.get_mpf_t();
  h.nr = v->m.rows(); h.nc = v->m.cols(); h.N = v->m.N;
  mp_bitcnt_t p = v->m.center.re.get_prec();
  if (v->m.center.im.get_prec() > p) p = v->m.center.im.get_prec();
  if (v->m.sz.re.get_prec() > p) p = v->m.sz.re.get_prec();
  if (v->m.sz.im.get_prec() > p) p = v->m.sz.im.get_prec();
  h.prec = p;
  return h;
}

template <typename F>
int guarded(nmv_view* v, F f) {
  if (!v) return NM_EINVAL;
  try {
    return f();
  } catch (const std::exception& e) {
    v->err = e.what();
    if (v->err.find("no CUDA device") != std::string::npos) return NM_ENODEV;
    if (v->err.find("exceed double range") != std::string::npos) return NM_ERANGE;
    return NM_ECUDA;
  } catch (...) {
    v->err = "unknown C++ exception";
    return NM_ECUDA;
  }
}
}  // namespace

extern "C" {

nmv_view* nmv_create(int nr, int nc) {
  if (nr < 1 || nc < 1) { g_err = "nmv_create: bad size"; return nullptr; }
  try {
    return new nmv_view(nr, nc);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void nmv_destroy(nmv_view* v) { delete v; }

const char* nmv_last_error(const nmv_view* v) { return v ? v->err.c_str() : g_err.c_str(); }

int nmv_set_view(nmv_view* v, int N, const char* sz_re, const char* sz_im, const char* c_re, const char* c_im,
                 double tol) {
  return guarded(v, [&]() {
    v->m.N = N;
    if (sz_re && sz_im) {
      v->m.sz.re = sz_re;
      v->m.sz.im = sz_im;
      v->m.zoom(1.0f);
    }
    if (c_re && c_im) {
      v->m.center.re = c_re;
      v->m.center.im = c_im;
    }
    v->m.error_tolerance = tol;
    return NM_OK;
  });
}

int nmv_set_floatexp(nmv_view* v, int force) {
  if (!v) return NM_EINVAL;
  v->m.force_floatexp = force < 0 ? 0 : (force > 2 ? 2 : force);
  return NM_OK;
}

int nmv_find_probe(nmv_view* v, int mode, int* row, int* col, int* length, int* n_exact) {
  return guarded(v, [&]() {
    const int saved = v->m.probe_search;
    v->m.probe_search = mode ? 1 : 0;
    int r = -1, c = -1, l = 0, ne = 0;
    try {
      v->m.findProbe(r, c, l, &ne);
    } catch (...) {
      v->m.probe_search = saved;
      throw;
    }
    v->m.probe_search = saved;
    if (row) *row = r;
    if (col) *col = c;
    if (length) *length = l;
    if (n_exact) *n_exact = ne;
    return NM_OK;
  });
}

int nmv_set_probe_search(nmv_view* v, int mode) {
  if (!v) return NM_EINVAL;
  v->m.probe_search = mode ? 1 : 0;
  return NM_OK;
}

int nmv_set_options(nmv_view* v, double glitch_tol, int max_secondary, int device, int host_threads) {
  if (!v) return NM_EINVAL;
  if (glitch_tol >= 0) v->m.glitch_tolerance = glitch_tol;
  if (max_secondary >= 0) v->m.max_secondary = max_secondary;
  if (device >= 0) v->m.device = device;
  if (host_threads >= 0) v->m.host_threads = host_threads;
  return NM_OK;
}

int nmv_get_n(const nmv_view* v) { return v ? v->m.N : NM_EINVAL; }
int nmv_rows(const nmv_view* v) { return v ? v->m.rows() : NM_EINVAL; }
int nmv_cols(const nmv_view* v) { return v ? v->m.cols() : NM_EINVAL; }
int nmv_use_hardware(nmv_view* v) { return v ? (v->m.useHardware() ? 1 : 0) : NM_EINVAL; }
int nmv_precision_bits(const nmv_view* v) { return v ? (int)v->m.center.re.get_prec() : NM_EINVAL; }

int nmv_precompute(nmv_view* v) { return guarded(v, [&]() { v->m.precompute(); return NM_OK; }); }
int nmv_compute_row(nmv_view* v, int r) {
  return guarded(v, [&]() {
    if (r < 0 || r >= v->m.rows()) { v->err = "row out of range"; return NM_EINVAL; }
    v->m.computeRow(r);
    return NM_OK;
  });
}
int nmv_render(nmv_view* v, nm_escape* out) {
  return guarded(v, [&]() {
    v->m.precompute();
    if (v->m.frameInfo().cancelled) return NM_ECANCELLED;   // nmv_cancel from another thread: no second attempt
    for (int r = 0; r < v->m.rows(); r++) v->m.computeRow(r);
    if (out) std::memcpy(out, v->m.g().values.data(), v->m.g().values.size() * sizeof(nm_escape));
    return NM_OK;
  });
}
int nmv_read_grid(const nmv_view* v, nm_escape* out) {
  if (!v || !out) return NM_EINVAL;
  std::memcpy(out, v->m.g().values.data(), v->m.g().values.size() * sizeof(nm_escape));
  return NM_OK;
}
int nmv_write_grid(nmv_view* v, const nm_escape* in) {
  if (!v || !in) return NM_EINVAL;
  std::memcpy(static_cast<void*>(v->m.g().values.data()), in, v->m.g().values.size() * sizeof(nm_escape));
  return NM_OK;
}
int nmv_at_sc(nmv_view* v, int r, int c, int sc, nm_escape* out) {
  if (!v || !out || sc < 1 || r < 0 || c < 0 || (r + 1) * sc > v->m.rows() || (c + 1) * sc > v->m.cols()) return NM_EINVAL;
  RenderGrid::EscapeValue e = v->m.at(r, c, sc);
  out->iterations = e.iterations;
  out->smoothing = e.smoothing;
  return NM_OK;
}
int nmv_scale(nmv_view* v, int sc, int up) {
  return guarded(v, [&]() {
    if (sc < 1) return NM_EINVAL;
    if (up) v->m.scaleUp(sc); else v->m.scaleDown(sc);
    return NM_OK;
  });
}
int nmv_zoom(nmv_view* v, float scale) { return guarded(v, [&]() { v->m.zoom(scale); return NM_OK; }); }
int nmv_translate(nmv_view* v, int dr, int dc, int sc) { return guarded(v, [&]() { v->m.translate(dr, dc, sc); return NM_OK; }); }
int nmv_zoom_at(nmv_view* v, float scale, int r, int c, int sc) {
  return guarded(v, [&]() { v->m.zoomAt(scale, r, c, sc); return NM_OK; });
}
int nmv_load_legacy(nmv_view* v, const char* fn) { return guarded(v, [&]() { v->m.loadLegacy(fn); return NM_OK; }); }
int nmv_save(nmv_view* v, const char* fn) { return guarded(v, [&]() { v->m.save(fn); return NM_OK; }); }

int nmv_view_string(const nmv_view* v, int which, char* buf, int cap) {
  if (!v || !buf || cap < 2 || which < 0 || which > 3) return NM_EINVAL;
  const mpf_class& f = which == 0 ? v->m.center.re : which == 1 ? v->m.center.im : which == 2 ? v->m.sz.re : v->m.sz.im;
  mp_exp_t e;
  char* s = mpf_get_str(nullptr, &e, 10, 0, f.get_mpf_t());
  int n = snprintf(buf, cap, "%s@%ld", s, (long)e);
  free(s);
  return n;
}

int nmv_frame_info_get(const nmv_view* v, nmv_frame_info* out) {
  if (!v || !out) return NM_EINVAL;
  fill_info(v->m.frameInfo(), out);
  return NM_OK;
}

int nmv_resolve(nmv_view* v, const uint8_t* pal_rgb, int n_pal, int sc, int smooth, uint8_t* rgb_out) {
  return guarded(v, [&]() { v->m.resolveRGB(pal_rgb, n_pal, sc, smooth != 0, rgb_out); return NM_OK; });
}

int nmv_host_tables(nmv_view* v, int row, int col, int* has_escape, int* probe_row, int* probe_col) {
  return guarded(v, [&]() {
    newman_b200::ViewHP h = hp_of(v);
    if (row < 0 || col < 0) {
      int len;
      newman_b200::find_probe(h, v->m.host_threads, row, col, len);
    }
    newman_b200::build_tables(h, row, col, v->tabs, v->m.host_threads);
    if (has_escape) *has_escape = v->tabs.has_escape ? 1 : 0;
    if (probe_row) *probe_row = row;
    if (probe_col) *probe_col = col;
    return v->tabs.M;
  });
}

int nmv_host_table(const nmv_view* v, int which, double* out) {
  if (!v || !out) return NM_EINVAL;
  const std::vector<double>* src = nullptr;
  switch (which) {
    case 0: src = &v->tabs.x_hi; break;
    case 1: src = &v->tabs.x_lo; break;
    case 2: src = &v->tabs.a; break;
    case 3: src = &v->tabs.b; break;
    case 4: src = &v->tabs.c; break;
    case 5: src = &v->tabs.eps_re; break;
    case 6: src = &v->tabs.eps_im; break;
    case 7: src = &v->tabs.a_m; break;
    case 8: src = &v->tabs.b_m; break;
    case 9: src = &v->tabs.c_m; break;
    case 10: src = &v->tabs.eps_re_m; break;
    case 11: src = &v->tabs.eps_im_m; break;
    default: return NM_EINVAL;
  }
  std::memcpy(out, src->data(), src->size() * sizeof(double));
  return (int)src->size();
}

int nmv_host_table_exp(const nmv_view* v, int which, int32_t* out) {
  if (!v || !out) return NM_EINVAL;
  const std::vector<int32_t>* src = which == 2 ? &v->tabs.a_e : which == 3 ? &v->tabs.b_e : which == 4 ? &v->tabs.c_e
                                    : which == 10 ? &v->tabs.eps_re_e : which == 11 ? &v->tabs.eps_im_e : nullptr;
  if (!src) return NM_EINVAL;
  std::memcpy(out, src->data(), src->size() * sizeof(int32_t));
  return (int)src->size();
}

int nmv_host_coords(nmv_view* v, double* c_re, double* c_im) {
  return guarded(v, [&]() {
    std::vector<double> a, b;
    newman_b200::pixel_coords(hp_of(v), a, b);
    std::memcpy(c_re, a.data(), a.size() * sizeof(double));
    std::memcpy(c_im, b.data(), b.size() * sizeof(double));
    return NM_OK;
  });
}

int nmv_host_cardioid(nmv_view* v, uint8_t* mask_or_null) {
  return guarded(v, [&]() {
    std::vector<uint8_t> mask;
    int mode = newman_b200::classify_cardioid(hp_of(v), v->m.host_threads, mask);
    if (mode == NM_CARDIOID_MASK && mask_or_null) std::memcpy(mask_or_null, mask.data(), mask.size());
    return mode;
  });
}

int nmv_host_in_cardioid(nmv_view* v, int r, int c) {
  return guarded(v, [&]() { return newman_b200::in_cardioid_pixel(hp_of(v), r, c) ? 1 : 0; });
}

// ---- multi-GPU render groups (multi_host.h; render_collective in mandelbrot_host.cpp) ----------------------------
int nmm_band_layout(int nr, int band_rows, int rank, int world, int32_t* rows_out) {
  if (nr < 0 || band_rows < 1 || world < 1 || rank < 0 || rank >= world || nr % band_rows) return NM_EINVAL;
  const int n_blocks = nr / band_rows;
  const int nb = newman_b200::RankLink::blocks_of(rank, world, n_blocks);
  if (rows_out)
    for (int r = 0; r < nb * band_rows; r++) rows_out[r] = ((r / band_rows) * world + rank) * band_rows + r % band_rows;
  return nb * band_rows;
}
int nmm_unique_id(uint8_t id[NMM_ID_BYTES]) {
  try { newman_b200::RankLink::unique_id(id); return NM_OK; }
  catch (const std::exception& e) { g_err = e.what(); return NM_ECUDA; }
}
int nmm_create(int device, int rank, int world, const uint8_t id[NMM_ID_BYTES], nmm_rank** out) {
  if (!out) return NM_EINVAL;
  *out = nullptr;
  try {
    nmm_rank* rk = new nmm_rank();
    rk->link.reset(new newman_b200::RankLink(device, rank, world, id));
    *out = rk;
    return NM_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return g_err.find("no CUDA device") != std::string::npos ? NM_ENODEV : NM_ECUDA;
  }
}
void nmm_destroy(nmm_rank* rk) { delete rk; }
const char* nmm_last_error(const nmm_rank* rk) { return rk ? rk->err.c_str() : g_err.c_str(); }
nm_ctx* nmm_ctx(nmm_rank* rk) { return rk ? rk->link->ctx : nullptr; }
double nmm_exchange_ms(const nmm_rank* rk) { return rk ? rk->link->exchange_ms() : 0.0; }
int nmm_render(nmm_rank* rk, nmv_view* view, int band_rows, nm_escape* out_raster, int return_mode, nmv_frame_info* info) {
  if (!rk || !view) return NM_EINVAL;
  try {
    newman_b200::FrameInfo f;
    newman_b200::render_collective(*rk->link, view->m, band_rows, out_raster, return_mode, f, nullptr);
    rk->nr = view->m.rows(); rk->nc = view->m.cols(); rk->N = view->m.N; rk->band = band_rows;
    if (info) fill_info(f, info);
    return NM_OK;
  } catch (const std::exception& e) { rk->err = e.what(); return NM_ECUDA; }
}
int nmm_resolve(nmm_rank* rk, const uint8_t* pal_rgb, int n_pal, int sc, int smooth, uint8_t* out_rgb, int return_mode) {
  if (!rk || !pal_rgb || sc < 1) return NM_EINVAL;
  try {
    newman_b200::RankLink& l = *rk->link;
    if (rk->band < 1 || rk->band % sc || rk->nc % sc) throw std::runtime_error("nmm_resolve: band_rows of the last nmm_render must be a multiple of sc");
    const int n_blocks = rk->nr / rk->band;
    const int nr_loc = newman_b200::RankLink::blocks_of(l.rank, l.world, n_blocks) * rk->band;
    const size_t block_bytes = (size_t)(rk->band / sc) * (rk->nc / sc) * 3;
    void* bd = l.band_buffer((size_t)(nr_loc > 0 ? nr_loc / sc : 1) * (rk->nc / sc) * 3);
    if (nr_loc > 0 && nm_resolve(l.ctx, pal_rgb, n_pal, rk->N, sc, smooth, (uint8_t*)bd) != NM_OK)
      throw std::runtime_error(std::string("nm_resolve: ") + nm_last_error(l.ctx));
    l.return_band(bd, block_bytes, n_blocks, out_rgb, return_mode);
    return NM_OK;
  } catch (const std::exception& e) { rk->err = e.what(); return NM_ECUDA; }
}

int nmv_set_exact(nmv_view* v, int on) {
  if (!v) return NM_EINVAL;
  v->m.exact = on ? 1 : 0;
  return NM_OK;
}

int nmv_cancel(nmv_view* v) {
  if (!v) return NM_EINVAL;
  v->m.cancel();   // no exception: atomics and nm_cancel only
  return NM_OK;
}

int nmv_set_devices(nmv_view* v, const int* devices, int n, int band_rows) {
  if (!v || n < 0 || (n > 0 && !devices)) return NM_EINVAL;
  v->m.devices.assign(devices, devices + n);
  if (band_rows > 0) v->m.band_rows = band_rows;
  return NM_OK;
}

int nmv_host_selfcheck(void) { return newman_b200::host_mpf_layout_ok() ? 1 : 0; }
}  // extern "C"
