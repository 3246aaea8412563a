//! `extern "C"` mirror of include/hijiki_b200.h and a `Renderer` with the reference's interface
//! (reference src/main.rs:1143-1424).  UNCOMPILED here (no Rust toolchain in this image); the same ABI
//! is exercised by every test through ctypes (hijiki_b200/_abi.py).
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

/// reference `ImageBlock`, src/main.rs:608-617 (40 bytes)
#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct ImageBlock {
    pub id: u32,
    pub seed: u32,
    pub origin: [u32; 2],
    pub dimension: [u32; 2],
    pub original_dimension: [u32; 2],
    pub sample_offset: [f32; 2],
}

#[repr(C)]
pub struct HjkArray {
    pub ptr: *const c_void,
    pub count: u64,
}

/// The 12 CompiledScene arrays in binding order (src/main.rs:314-327).
#[repr(C)]
pub struct HjkScene {
    pub scene: HjkArray,
    pub bvh: HjkArray,
    pub spheres: HjkArray,
    pub quads: HjkArray,
    pub triangles: HjkArray,
    pub vertices: HjkArray,
    pub materials: HjkArray,
    pub emitters: HjkArray,
    pub diffuse: HjkArray,
    pub diffusecb: HjkArray,
    pub dielectric: HjkArray,
    pub emissive: HjkArray,
}

#[repr(C)]
pub struct HjkParams {
    pub max_bounces: u32,  // render.glsl:92 (1000)
    pub rr_start: u32,     // render.glsl:137 (3)
    pub recon_radius: u32, // src/main.rs:1284 (2)
    pub recon_stddev: f32, // src/main.rs:1285 (0.5)
    pub eps: f32,          // math.glsl:2 (1e-4)
    pub flags: u32,
}

impl Default for HjkParams {
    fn default() -> Self {
        HjkParams { max_bounces: 1000, rr_start: 3, recon_radius: 2, recon_stddev: 0.5, eps: 1e-4, flags: 0 }
    }
}

#[repr(C)]
#[derive(Default, Debug)]
pub struct HjkStats {
    pub n_paths: u64,
    pub n_extension_rays: u64,
    pub n_shadow_rays: u64,
    pub ms_total: f32,
    pub kernel_ms: [f32; 8],
    pub n_launches: u64,
}

pub enum HjkContext {}

extern "C" {
    pub fn hjk_create(device_ids: *const c_int, n_devices: c_int, out_ctx: *mut *mut HjkContext) -> c_int;
    pub fn hjk_destroy(ctx: *mut HjkContext) -> c_int;
    pub fn hjk_last_error(ctx: *const HjkContext) -> *const c_char;
    pub fn hjk_scene_upload(ctx: *mut HjkContext, scene: *const HjkScene) -> c_int;
    pub fn hjk_frame_begin(ctx: *mut HjkContext, width: u32, height: u32) -> c_int;
    pub fn hjk_render(ctx: *mut HjkContext, blocks: *const ImageBlock, n_blocks: u64, params: *const HjkParams,
                      stats: *mut HjkStats) -> c_int;
    pub fn hjk_blocks_upload(ctx: *mut HjkContext, blocks: *const ImageBlock, n_blocks: u64, out_handle: *mut u64) -> c_int;
    pub fn hjk_render_resident(ctx: *mut HjkContext, handle: u64, first_block: u64, n_blocks: u64,
                               params: *const HjkParams, stats: *mut HjkStats) -> c_int;
    pub fn hjk_blocks_free(ctx: *mut HjkContext, handle: u64) -> c_int;
    pub fn hjk_readback(ctx: *mut HjkContext, rgba: *mut f32, pitch_bytes: u64, normalise: c_int) -> c_int;
    pub fn hjk_readback_root(ctx: *mut HjkContext, root: c_int, rgba: *mut f32, pitch_bytes: u64, normalise: c_int) -> c_int;
    pub fn hjk_readback_begin(ctx: *mut HjkContext, root: c_int, rgba: *mut f32, pitch_bytes: u64, normalise: c_int) -> c_int;
    pub fn hjk_readback_wait(ctx: *mut HjkContext) -> c_int;
    pub fn hjk_read_features(ctx: *mut HjkContext, root: c_int, normal_depth: *mut f32, pitch_bytes: u64) -> c_int;
    pub fn hjk_read_intermediate(ctx: *mut HjkContext, layer: c_int, rgba: *mut f32) -> c_int;
    pub fn hjk_trace_first_hit(ctx: *mut HjkContext, rays: *const c_void, n_rays: u64, any_hit: c_int,
                               shape_id: *mut i32, t: *mut f32, uv: *mut f32) -> c_int;
    pub fn hjk_denoise_pass(ctx: *mut HjkContext, radiance: *const f32, normal_depth: *const f32, albedo: *const f32,
                            blocks: *const ImageBlock, n_blocks: u64, params: *const HjkParams) -> c_int;
    pub fn hjk_comm_unique_id(out_id128: *mut c_void) -> c_int;
    pub fn hjk_comm_init(ctx: *mut HjkContext, id128: *const c_void, rank: c_int, n_ranks: c_int) -> c_int;
    pub fn hjk_reduce_frame(ctx: *mut HjkContext, root: c_int, out_ms: *mut f32) -> c_int;
    pub fn hjk_allreduce_accumulator(ctx: *mut HjkContext, out_ms: *mut f32) -> c_int;
    pub fn hjk_accumulator_device_ptr(ctx: *mut HjkContext, out_ptr: *mut u64, out_n_floats: *mut u64) -> c_int;
    pub fn hjk_synchronize(ctx: *mut HjkContext) -> c_int;
    pub fn hjk_set_stream(ctx: *mut HjkContext, cuda_stream: *mut c_void) -> c_int;
    pub fn hjk_set_profiling(ctx: *mut HjkContext, enabled: c_int) -> c_int;
    pub fn hjk_set_option(ctx: *mut HjkContext, key: *const c_char, value: i64) -> c_int;
    pub fn hjk_get_info(ctx: *mut HjkContext, key: *const c_char, out_value: *mut i64) -> c_int;
    pub fn hjk_version() -> *const c_char;
}

fn check(ctx: *const HjkContext, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(hjk_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("hijiki_b200 error {}: {}", rc, msg);
    }
}

pub fn view<T>(v: &[T]) -> HjkArray {
    HjkArray { ptr: v.as_ptr() as *const c_void, count: v.len() as u64 }
}

/// Stand-in for the GPU half of the reference's `Renderer` (src/main.rs:1143-1424).
/// `scene` is built from the host's `CompiledScene` with `view(&compiled.spheres)` etc.
pub struct Renderer {
    ctx: *mut HjkContext,
    blocks: Vec<ImageBlock>,
    width: u32,
    height: u32,
    pub params: HjkParams,
}

impl Renderer {
    /// `Renderer::new(scene, generator, present_interval, use_bvh)`: uploads the scene (copies) and
    /// zeroes the accumulator.  `rank`/`n_ranks` select the sample passes of this process.
    pub fn new(scene: &HjkScene, blocks: Vec<ImageBlock>, width: u32, height: u32, device: i32) -> Self {
        let mut ctx = std::ptr::null_mut();
        unsafe {
            check(std::ptr::null(), hjk_create([device].as_ptr(), 1, &mut ctx));
            check(ctx, hjk_scene_upload(ctx, scene));
            check(ctx, hjk_frame_begin(ctx, width, height));
        }
        Renderer { ctx, blocks, width, height, params: HjkParams::default() }
    }

    /// `Renderer::render` (src/main.rs:1316-1355): every block integrated and reconstructed.
    pub fn render(&mut self) -> HjkStats {
        let mut st = HjkStats::default();
        unsafe {
            check(self.ctx, hjk_render(self.ctx, self.blocks.as_ptr(), self.blocks.len() as u64, &self.params, &mut st));
        }
        st
    }

    /// `save_image`'s readback + divide (src/main.rs:1357-1400): (r/w, g/w, b/w, w) per texel.
    pub fn image(&mut self) -> Vec<f32> {
        let mut rgba = vec![0f32; (self.width * self.height * 4) as usize];
        unsafe { check(self.ctx, hjk_readback(self.ctx, rgba.as_mut_ptr(), self.width as u64 * 16, 1)); }
        rgba
    }
}

impl Drop for Renderer {
    fn drop(&mut self) {
        unsafe { hjk_destroy(self.ctx); }
    }
}
