//! Raw bindings of `include/jxlb200.h` (ABI version 2).  Field order and types mirror the C header exactly;
//! `tests/test_rust_binding_decls.py` in the repository checks this file against it.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const JXLB200_ABI_VERSION: c_int = 3;

#[repr(C)]
pub struct jxlb200_ctx {
    _private: [u8; 0],
}

/// 8-bit sRGB, interleaved RGB, row-major; `stride` is bytes per row (>= 3 * width).
#[repr(C)]
pub struct jxlb200_image {
    pub pixels: *const u8,
    pub width: u32,
    pub height: u32,
    pub stride: usize,
}

/// cjxl's `--distance` / `--effort` (docker_manager.rs:125-127) plus the proposal toggle.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct jxlb200_params {
    pub distance: f32,
    pub effort: u32,
    pub proposal: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct jxlb200_stats {
    pub codestream_bytes: u64,
    pub bpp: f64,
    pub width: u32,
    pub height: u32,
    pub num_groups: u32,
    pub num_dc_groups: u32,
    pub global_scale: u32,
    pub quant_dc: u32,
    pub num_tokens: u64,
    pub num_clusters: u32,
    pub acs_histogram: [u32; 27],
    pub stage_ms: [f32; 16],
    pub total_ms: f32,
    pub kernel_launches: u32,
    pub quality_valid: u32,
    pub sse: [u64; 3],
    pub psnr: f64,
}

pub const JXLB200_PROPOSAL_NONE: u32 = 0; // unpatched libjxl ("main", context.rs:17)
pub const JXLB200_PROPOSAL_PARTITIONING: u32 = 1; // proposals/homogeneity-partitioning.diff
pub const JXLB200_PROPOSAL_FACTORED_ENTROPY: u32 = 2; // proposals/homogeneity-factored-entropy.diff
pub const JXLB200_PROPOSAL_COMBINED: u32 = 3; // proposals/combined.diff

pub const JXLB200_FLAG_FIXED_DCT8: u32 = 1;
pub const JXLB200_FLAG_UNIFORM_QF: u32 = 2;
pub const JXLB200_FLAG_QUALITY: u32 = 4;
pub const JXLB200_FLAG_FORCED_ACS: u32 = 8;
pub const JXLB200_FLAG_GABORISH: u32 = 16;
pub const JXLB200_FLAG_CFL: u32 = 32;

extern "C" {
    pub fn jxlb200_abi_version() -> c_int;
    pub fn jxlb200_create(device: c_int) -> *mut jxlb200_ctx;
    pub fn jxlb200_destroy(ctx: *mut jxlb200_ctx);
    pub fn jxlb200_last_error(ctx: *const jxlb200_ctx) -> *const c_char;
    pub fn jxlb200_set_pipelines(ctx: *mut jxlb200_ctx, n: c_int) -> c_int;
    pub fn jxlb200_encode(
        ctx: *mut jxlb200_ctx,
        image: *const jxlb200_image,
        params: *const jxlb200_params,
        out: *mut *mut u8,
        out_len: *mut usize,
        stats: *mut jxlb200_stats,
    ) -> c_int;
    pub fn jxlb200_encode_batch(
        ctx: *mut jxlb200_ctx,
        images: *const jxlb200_image,
        params: *const jxlb200_params,
        n: usize,
        outs: *mut *mut u8,
        out_lens: *mut usize,
        stats: *mut jxlb200_stats,
    ) -> c_int;
    pub fn jxlb200_encode_device(
        ctx: *mut jxlb200_ctx,
        d_pixels: *const u8,
        width: u32,
        height: u32,
        stride: usize,
        params: *const jxlb200_params,
        stats: *mut jxlb200_stats,
    ) -> c_int;
    pub fn jxlb200_encode_batch_device(
        ctx: *mut jxlb200_ctx,
        d_pixels: *const *const u8,
        widths: *const u32,
        heights: *const u32,
        strides: *const usize,
        params: *const jxlb200_params,
        n: usize,
        stats: *mut jxlb200_stats,
        device_ms: *mut f32,
    ) -> c_int;
    pub fn jxlb200_fetch(ctx: *mut jxlb200_ctx, out: *mut *mut u8, out_len: *mut usize) -> c_int;
    pub fn jxlb200_free(buf: *mut c_void);
    pub fn jxlb200_dump(ctx: *mut jxlb200_ctx, stage: c_int, dst: *mut c_void, cap: usize) -> i64;
    pub fn jxlb200_debug_set_strategy_map(ctx: *mut jxlb200_ctx, acs: *const u8, bxs: u32, bys: u32) -> c_int;
    pub fn jxlb200_debug_homogeneity(
        ctx: *mut jxlb200_ctx,
        x: *const f32,
        y: *const f32,
        b: *const f32,
        stride: u32,
        ysize: u32,
        distance: f32,
        out: *mut f32,
    ) -> c_int;
    pub fn jxlb200_dims(width: u32, height: u32, dims: *mut i32);
}
