//! Raw bindings to include/blbm.h — one `extern "C"` item per entry point, same order as the header.
//! UNVERIFIED SOURCE: this image has no cargo/rustc; the ABI itself is exercised through Python ctypes
//! (lbm_b200/lbm.py, tests/test_abi.py checks the prototype table against the header).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_float, c_int, c_void};

#[repr(C)]
pub struct blbm_t {
    _private: [u8; 0],
}

pub const BLBM_OK: c_int = 0;
pub const BLBM_EINVAL: c_int = -1;
pub const BLBM_ENOMEM: c_int = -2;
pub const BLBM_ECUDA: c_int = -3;
pub const BLBM_ENOGPU: c_int = -4;
pub const BLBM_ESTATE: c_int = -5;
pub const BLBM_EPEER: c_int = -6;
pub const BLBM_PEER_HANDLE_BYTES: usize = 512;

// blbm_tune (include/blbm.h): launch-shape knobs for A/B measurement; results never depend on them
pub const BLBM_TUNE_VEC4_BLOCK_ROWS: c_int = 0;
pub const BLBM_TUNE_VEC4_DENSE: c_int = 4;
pub const BLBM_TUNE_CUDA_GRAPHS: c_int = 5;
pub const BLBM_TUNE_VEC4_PACKED: c_int = 6;
pub const BLBM_TUNE_VEC4_INDEX32: c_int = 7;
pub const BLBM_TUNE_LINK_IN_KERNEL: c_int = 8;
pub const BLBM_TUNE_PDL: c_int = 9;

extern "C" {
    pub fn blbm_last_error() -> *const c_char;
    pub fn blbm_abi_version() -> c_int;
    pub fn blbm_device_count() -> c_int;
    pub fn blbm_create(w: u32, h: u32, omega: c_float, inflow_ux: c_float, device: c_int, out: *mut *mut blbm_t) -> c_int;
    pub fn blbm_create_slab(w: u32, h_global: u64, row_begin: u64, row_end: u64, omega: c_float, inflow_ux: c_float,
                            device: c_int, out: *mut *mut blbm_t) -> c_int;
    pub fn blbm_create_group(w: u32, h: u64, omega: c_float, inflow_ux: c_float, devices: *const c_int, ndev: c_int,
                             out: *mut *mut blbm_t) -> c_int;
    pub fn blbm_group_size(h: *const blbm_t) -> c_int;
    pub fn blbm_group_slab(h: *mut blbm_t, index: c_int, slab: *mut *mut blbm_t) -> c_int;
    pub fn blbm_destroy(h: *mut blbm_t) -> c_int;
    pub fn blbm_iterate(h: *mut blbm_t, n: u32) -> c_int;
    pub fn blbm_advance(h: *mut blbm_t, n: u32) -> c_int;
    pub fn blbm_iterate_timed(h: *mut blbm_t, n: u32, elapsed_ms: *mut c_float) -> c_int;
    pub fn blbm_timer_start(h: *mut blbm_t) -> c_int;
    pub fn blbm_timer_stop(h: *mut blbm_t, elapsed_ms: *mut c_float) -> c_int;
    pub fn blbm_collide(h: *mut blbm_t) -> c_int;
    pub fn blbm_stream(h: *mut blbm_t) -> c_int;
    pub fn blbm_set_summary(h: *mut blbm_t, stat: c_int) -> c_int;
    pub fn blbm_rerender(h: *mut blbm_t) -> c_int;
    pub fn blbm_compute_summary(h: *mut blbm_t, stat: c_int) -> c_int;
    pub fn blbm_set_omega(h: *mut blbm_t, omega: c_float) -> c_int;
    pub fn blbm_reset_to_equilibrium(h: *mut blbm_t) -> c_int;
    pub fn blbm_custom_speed(h: *mut blbm_t, ux: c_float) -> c_int;
    pub fn blbm_single_cell(h: *mut blbm_t, index: u32) -> c_int;
    pub fn blbm_draw_points(h: *mut blbm_t, loc_val_pairs: *const u32, npairs: usize) -> c_int;
    pub fn blbm_draw_points64(h: *mut blbm_t, loc_val_pairs: *const u64, npairs: usize) -> c_int;
    pub fn blbm_reset_barrier(h: *mut blbm_t) -> c_int;
    pub fn blbm_write_barrier_rows(h: *mut blbm_t, row_begin: u64, nrows: u64, mask: *const u8) -> c_int;
    pub fn blbm_rasterize_line(x1: i64, y1: i64, x2: i64, y2: i64, xdim: i64, ydim: i64, erase: c_int, xy: *mut i64,
                               capacity: usize, count: *mut usize) -> c_int;
    pub fn blbm_preset_lines(preset: c_int, xdim: i64, ydim: i64, xyxy: *mut i64, capacity: usize, count: *mut usize) -> c_int;
    pub fn blbm_draw_line(h: *mut blbm_t, x1: i64, y1: i64, x2: i64, y2: i64) -> c_int;
    pub fn blbm_erase_line(h: *mut blbm_t, x1: i64, y1: i64, x2: i64, y2: i64) -> c_int;
    pub fn blbm_curl_barrier(h: *mut blbm_t) -> c_int;
    pub fn blbm_chaos_barrier(h: *mut blbm_t) -> c_int;
    pub fn blbm_welcome_barrier(h: *mut blbm_t) -> c_int;
    pub fn blbm_get_compute_num(h: *const blbm_t) -> u64;
    pub fn blbm_get_frame_num(h: *const blbm_t) -> u64;
    pub fn blbm_read_population(h: *mut blbm_t, buffer: c_int, k: c_int, dst: *mut c_float) -> c_int;
    pub fn blbm_write_population(h: *mut blbm_t, buffer: c_int, k: c_int, src: *const c_float) -> c_int;
    pub fn blbm_read_moments(h: *mut blbm_t, mx: *mut c_float, my: *mut c_float, rho: *mut c_float) -> c_int;
    pub fn blbm_read_output(h: *mut blbm_t, dst: *mut c_float) -> c_int;
    pub fn blbm_read_barrier(h: *mut blbm_t, dst: *mut u32) -> c_int;
    pub fn blbm_read_cell_class(h: *mut blbm_t, dst: *mut u16) -> c_int;
    pub fn blbm_color_map(h: *mut blbm_t, map: c_int) -> c_int;
    pub fn blbm_read_colors(h: *mut blbm_t, rgb: *mut c_float) -> c_int;
    pub fn blbm_read_output_async(h: *mut blbm_t, pinned_dst: *mut c_float) -> c_int;
    pub fn blbm_synchronize(h: *mut blbm_t) -> c_int;
    pub fn blbm_reduce_moments(h: *mut blbm_t, sum_rho: *mut c_double, sum_mx: *mut c_double, sum_my: *mut c_double,
                               max_abs_output: *mut c_float) -> c_int;
    pub fn blbm_get_geometry(h: *const blbm_t, w: *mut u32, h_global: *mut u64, row_begin: *mut u64, row_end: *mut u64,
                             device: *mut c_int) -> c_int;
    pub fn blbm_link_local(upper: *mut blbm_t, lower: *mut blbm_t) -> c_int;
    pub fn blbm_export_peer(h: *mut blbm_t, blob: *mut c_void) -> c_int;
    pub fn blbm_link_peer(h: *mut blbm_t, side: c_int, blob: *const c_void) -> c_int;
    pub fn blbm_exchange_halos(h: *mut blbm_t) -> c_int;
    pub fn blbm_set_kernel(h: *mut blbm_t, kernel: c_int) -> c_int;
    pub fn blbm_get_kernel(h: *const blbm_t) -> c_int;
    pub fn blbm_set_tuning(h: *mut blbm_t, knob: c_int, value: c_int) -> c_int;
    pub fn blbm_set_lazy_barriers(h: *mut blbm_t, mode: c_int) -> c_int;
    pub fn blbm_get_lazy_barriers_active(h: *const blbm_t) -> c_int;
    pub fn blbm_get_launch_count(h: *const blbm_t) -> u64;
    pub fn blbm_get_device_bytes(h: *const blbm_t) -> u64;
}
