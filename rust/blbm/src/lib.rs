//! Drop-in for `lbm-wgpu/src/lbm.rs`'s `pub struct LBM` on top of libblbm.so.
//!
//! Same method names and argument meaning; `&Driver` survives as a zero-size `Driver` so that call sites
//! in lib.rs (`:46,65,122-197`) compile unchanged.  `render`/`color_map` stay with the caller (they need
//! a surface); `read_*` are new.  UNVERIFIED SOURCE (no cargo/rustc in the image).
use blbm_sys as sys;
use std::ffi::CStr;

#[derive(PartialEq, Clone, Copy, Debug)]
pub enum SummaryStat {
    Curl = 0,
    Ux = 1,
    Uy = 2,
    Rho = 3,
    Speed = 4,
}

/// Stand-in for lbm-wgpu's `Driver` (driver.rs:1-7): the CUDA devices and streams live inside the handle.
/// One device: the whole lattice on it.  Several: `LBM::new` cuts the lattice into y-slabs, one per listed
/// device, behind the same `LBM` value (blbm_create_group), so a 65536 x 65536 lattice is constructed with
/// `Driver { devices: (0..8).collect() }` and stepped, painted and read exactly like a small one.
#[derive(Clone)]
pub struct Driver {
    pub devices: Vec<i32>,
}

impl Default for Driver {
    fn default() -> Driver {
        Driver { devices: vec![0] }
    }
}

/// `trait Shape` of barrier_shapes/mod.rs:11-19, reduced to what draw_shape needs.
pub trait Shape {
    fn get_points(&self) -> &std::collections::HashSet<(isize, isize, bool)>;
}

pub struct LBM {
    h: *mut sys::blbm_t,
    x: u32,
    y: u32,
    pub compute_step: usize,
}

fn check(rc: i32) {
    if rc != sys::BLBM_OK {
        // the reference panics on every failure (expect/unwrap); keep that contract
        let msg = unsafe { CStr::from_ptr(sys::blbm_last_error()) }.to_string_lossy().into_owned();
        panic!("blbm error {}: {}", rc, msg);
    }
}

impl LBM {
    /// lbm.rs:726
    pub fn new(driver: &Driver, omega: f32, x: u32, y: u32) -> LBM {
        let mut h = std::ptr::null_mut();
        check(unsafe {
            sys::blbm_create_group(x, y as u64, omega, 0.1, driver.devices.as_ptr(), driver.devices.len() as i32, &mut h)
        });
        LBM { h, x, y, compute_step: 0 }
    }
    /// lbm.rs:1065 (minus colour map and render)
    pub fn iterate(&mut self, _driver: &Driver, compute_steps: usize) {
        check(unsafe { sys::blbm_iterate(self.h, compute_steps as u32) });
        self.compute_step = unsafe { sys::blbm_get_compute_num(self.h) } as usize;
    }
    /// lbm.rs:1118
    pub fn collide(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_collide(self.h) });
    }
    /// lbm.rs:1127
    pub fn stream(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_stream(self.h) });
    }
    /// lbm.rs:1104 (summary only)
    pub fn rerender(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_rerender(self.h) });
    }
    /// lbm.rs:1061
    pub fn set_summary(&mut self, stat: SummaryStat) {
        check(unsafe { sys::blbm_set_summary(self.h, stat as i32) });
    }
    /// lbm.rs:1337 + merge_shapes.rs:12-22
    pub fn draw_shape(&mut self, _driver: &Driver, shape: &dyn Shape) {
        let pts: Vec<u32> = shape
            .get_points()
            .iter()
            .flat_map(|p| vec![p.0 as u32 + p.1 as u32 * self.x, if p.2 { 1 } else { 0 }])
            .collect();
        check(unsafe { sys::blbm_draw_points(self.h, pts.as_ptr(), pts.len() / 2) });
    }
    /// lbm.rs:1362
    pub fn reset_barrier(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_reset_barrier(self.h) });
    }
    /// lbm.rs:1358
    pub fn update_omega_buffer(&mut self, _driver: &Driver, omega: f32) {
        check(unsafe { sys::blbm_set_omega(self.h, omega) });
    }
    /// lbm.rs:1076
    pub fn reset_to_equilibrium(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_reset_to_equilibrium(self.h) });
        self.compute_step = 0;
    }
    /// lbm.rs:1090
    pub fn custom_speed(&mut self, _driver: &Driver, ux: f32) {
        check(unsafe { sys::blbm_custom_speed(self.h, ux) });
        self.compute_step = 0;
    }
    /// lbm.rs:1502
    pub fn single_cell(&mut self, _driver: &Driver, index: usize) {
        check(unsafe { sys::blbm_single_cell(self.h, index as u32) });
        self.compute_step = 0;
    }
    /// lbm.rs:1367 / :1372 / :1388 — the barrier presets (thick lines through the library's rasteriser)
    pub fn curl_barrier(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_curl_barrier(self.h) });
    }
    pub fn chaos_barrier(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_chaos_barrier(self.h) });
    }
    pub fn welcome_barrier(&mut self, _driver: &Driver) {
        check(unsafe { sys::blbm_welcome_barrier(self.h) });
    }
    /// lbm.rs:1299 — colour-map the output field on the device (Inferno 0, Viridis 1, Jet 2; lbm.rs:18-24)
    pub fn color_map(&mut self, map: i32) {
        check(unsafe { sys::blbm_color_map(self.h, map) });
    }
    pub fn read_colors(&mut self) -> Vec<f32> {
        let mut v = vec![0f32; self.x as usize * self.y as usize * 3];
        check(unsafe { sys::blbm_read_colors(self.h, v.as_mut_ptr()) });
        v
    }
    /// lbm.rs:1166
    pub fn get_frame_num(&self) -> usize {
        unsafe { sys::blbm_get_frame_num(self.h) as usize }
    }
    /// lbm.rs:1170
    pub fn get_compute_num(&self) -> usize {
        unsafe { sys::blbm_get_compute_num(self.h) as usize }
    }

    // ---- new: read back macroscopic fields ----
    pub fn read_output(&mut self) -> Vec<f32> {
        let mut v = vec![0f32; self.x as usize * self.y as usize];
        check(unsafe { sys::blbm_read_output(self.h, v.as_mut_ptr()) });
        v
    }
    pub fn read_moments(&mut self) -> (Vec<f32>, Vec<f32>, Vec<f32>) {
        let n = self.x as usize * self.y as usize;
        let (mut a, mut b, mut c) = (vec![0f32; n], vec![0f32; n], vec![0f32; n]);
        check(unsafe { sys::blbm_read_moments(self.h, a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr()) });
        (a, b, c)
    }
    pub fn read_population(&mut self, k: usize) -> Vec<f32> {
        let mut v = vec![0f32; self.x as usize * self.y as usize];
        check(unsafe { sys::blbm_read_population(self.h, -1, k as i32, v.as_mut_ptr()) });
        v
    }
}

impl Drop for LBM {
    fn drop(&mut self) {
        unsafe { sys::blbm_destroy(self.h) };
    }
}
