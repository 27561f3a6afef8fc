//! Raw bindings of include/fdf.h (the C ABI of libfdf_cuda.so).  Hand-written, one item per declaration.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct fdf_ctx {
    _private: [u8; 0],
}

#[repr(C)]
pub struct fdf_pipe {
    _private: [u8; 0],
}

/// `fdf_point` == `crate::Point` (`#[repr(C)] { x: u32, y: u32 }`).
pub type fdf_point = crate::Point;

pub const FDF_OK: c_int = 0;
pub const FDF_ERR_INVALID_COUNT: c_int = 1;
pub const FDF_ERR_CAPACITY: c_int = 4;
pub const FDF_ERR_BUSY: c_int = 8;

extern "C" {
    pub fn fdf_create(device: c_int, out_ctx: *mut *mut fdf_ctx) -> c_int;
    pub fn fdf_destroy(ctx: *mut fdf_ctx);
    pub fn fdf_detect(
        ctx: *mut fdf_ctx,
        img: *const u8,
        w: u32,
        h: u32,
        pitch: u32,
        threshold: u8,
        count: u8,
        nms: u8,
        out: *mut fdf_point,
        cap: usize,
        n_out: *mut usize,
    ) -> c_int;
    pub fn fdf_detect_batch(
        ctx: *mut fdf_ctx,
        frames: *const u8,
        n_frames: u32,
        w: u32,
        h: u32,
        pitch: u32,
        frame_stride: u64,
        threshold: u8,
        count: u8,
        nms: u8,
        out: *mut fdf_point,
        cap: usize,
        offsets: *mut u64,
    ) -> c_int;
    pub fn fdf_detect_device(
        ctx: *mut fdf_ctx,
        d_frames: *const u8,
        n_frames: u32,
        w: u32,
        h: u32,
        pitch: u32,
        frame_stride: u64,
        threshold: u8,
        count: u8,
        nms: u8,
        d_out: *mut fdf_point,
        cap: usize,
        d_offsets: *mut u64,
        stream: *mut c_void,
    ) -> c_int;
    /// main.rs:53-58 `to_rgb8()` + `to_luma8()` on device memory (include/fdf.h: fdf_rgb8_to_luma8_device)
    pub fn fdf_rgb8_to_luma8_device(
        ctx: *mut fdf_ctx,
        d_rgb: *const u8,
        n_frames: u32,
        w: u32,
        h: u32,
        rgb_pitch: u32,
        rgb_frame_stride: u64,
        d_luma: *mut u8,
        luma_pitch: u32,
        luma_frame_stride: u64,
        stream: *mut core::ffi::c_void,
    ) -> c_int;
    /// util.rs:5-41 `Rgb8ToLuma16View` + `to_grey`: (r + g + b) / 3 (include/fdf.h: fdf_rgb8_to_grey_sum3_device)
    pub fn fdf_rgb8_to_grey_sum3_device(
        ctx: *mut fdf_ctx,
        d_rgb: *const u8,
        n_frames: u32,
        w: u32,
        h: u32,
        rgb_pitch: u32,
        rgb_frame_stride: u64,
        d_grey: *mut u8,
        grey_pitch: u32,
        grey_frame_stride: u64,
        stream: *mut core::ffi::c_void,
    ) -> c_int;
    /// main.rs:53-67 in one call on a host RGB8 image (include/fdf.h: fdf_detect_rgb8)
    pub fn fdf_detect_rgb8(
        ctx: *mut fdf_ctx,
        rgb: *const u8,
        w: u32,
        h: u32,
        rgb_pitch: u32,
        threshold: u8,
        count: u8,
        nms: u8,
        out: *mut fdf_point,
        cap: usize,
        n_out: *mut usize,
    ) -> c_int;
    // streaming form of fdf_detect: up to `depth` images in flight, results first in, first out
    pub fn fdf_pipe_create(
        ctx: *mut fdf_ctx,
        depth: u32,
        max_w: u32,
        max_h: u32,
        cap: usize,
        out_pipe: *mut *mut fdf_pipe,
    ) -> c_int;
    pub fn fdf_pipe_destroy(pipe: *mut fdf_pipe);
    pub fn fdf_pipe_submit(
        pipe: *mut fdf_pipe,
        img: *const u8,
        w: u32,
        h: u32,
        pitch: u32,
        threshold: u8,
        count: u8,
        nms: u8,
    ) -> c_int;
    pub fn fdf_pipe_collect(pipe: *mut fdf_pipe, out: *mut fdf_point, cap: usize, n_out: *mut usize) -> c_int;
    pub fn fdf_pipe_in_flight(pipe: *const fdf_pipe) -> u32;
    // sharded batches (one process per GPU): the exchange step and the peer-mapped result buffer
    pub fn fdf_shard_push(
        ctx: *mut fdf_ctx,
        d_all_offsets: *const u64,
        block: u32,
        n_ranks: u32,
        rank: u32,
        total_frames: u32,
        d_points: *const fdf_point,
        d_result: *mut fdf_point,
        cap_total: usize,
        d_global_offsets: *mut u64,
        stream: *mut c_void,
    ) -> c_int;
    pub fn fdf_shared_alloc(ctx: *mut fdf_ctx, bytes: usize, d_ptr: *mut *mut c_void, handle: *mut u8) -> c_int;
    pub fn fdf_shared_open(ctx: *mut fdf_ctx, handle: *const u8, d_ptr: *mut *mut c_void) -> c_int;
    pub fn fdf_shared_close(ctx: *mut fdf_ctx, d_ptr: *mut c_void) -> c_int;
    pub fn fdf_set_tuning(ctx: *mut fdf_ctx, strip_rows: c_int, sub_batch_mb: u32) -> c_int;
    pub fn fdf_set_item_parts(ctx: *mut fdf_ctx, parts: u32) -> c_int;
    pub fn fdf_set_idle_sms(ctx: *mut fdf_ctx, sm_stride: u32) -> c_int;
    pub fn fdf_last_error(ctx: *const fdf_ctx) -> *const c_char;
    pub fn fdf_status_string(status: c_int) -> *const c_char;
}
