/*!
    FAST feature detector (Rosten & Drummond, 2006) -- B200 edition.

    Drop-in for the detection path of `iwanders/feature_detector_fast`: `Point`, `NonMaximalSuppression`,
    `Config`, `Config::detect` and `detect` keep the reference's signatures and derives (reference
    `src/lib.rs:15-64`); the body that used to forward to the AVX2 module `fast_simd::detector` now calls the
    CUDA library through the C ABI of `include/fdf.h`.  Keypoint sets, order and the panics for `count`
    outside 9..=16 are those of the reference.

    NOT COMPILED in the build image (no Rust toolchain there): this file is the binding a maintainer
    adds; everything it calls is tested through the same C ABI from C++ and Python.
*/
pub mod ffi;

// The reference's scalar specification and its drawing helpers stay in the crate exactly as they are
// (reference `src/lib.rs:9-10`): `tests/compare.rs:1,49`, `benches/benchmark.rs:2` and `src/main.rs:1` import them.
// The two files are the maintainer's own `src/opencv_compat.rs` and `src/util.rs`, unchanged; they are not
// duplicated in this repository (this directory is an overlay on the reference tree).
pub mod opencv_compat;
pub mod util;

/// Stands where the AVX2 module stood (reference `src/lib.rs:12-13`, `src/fast_simd.rs:847-859`), so that
/// `tests/compare.rs:45` and `benches/benchmark.rs:25,36,47`, which call `fast_simd::detector` directly, build and
/// run against the CUDA path unchanged.  No `target_feature = "avx2"` gate any more: the path needs a B200, not AVX2.
pub mod fast_simd {
    use crate::{Config, Point};

    /// Same signature and results as the reference's `fast_simd::detector` (`src/fast_simd.rs:847`).
    pub fn detector(image: &image::GrayImage, config: &Config) -> Vec<Point> {
        crate::detect(image, config)
    }
}

use std::cell::RefCell;
use std::ffi::CStr;

#[repr(C)]
#[derive(Copy, Debug, Clone, Eq, PartialEq, Hash, Default)]
/// A feature point at an image position.
pub struct Point {
    pub x: u32,
    pub y: u32,
}

/// Modes of non maximal suppression (see the reference for the description of each).
#[derive(Debug, Copy, Clone, Eq, PartialEq, Ord, PartialOrd, Hash)]
pub enum NonMaximalSuppression {
    Off,
    MaxThreshold,
    SumAbsolute,
}

#[derive(Copy, Clone, Debug, Eq, PartialEq, Ord, PartialOrd, Hash)]
/// Configuration struct for the FAST feature detector.
pub struct Config {
    pub threshold: u8,
    /// Allowed values are count >= 9 && count <= 16.
    pub count: u8,
    pub non_maximal_supression: NonMaximalSuppression,
}

impl Config {
    /// Method access to run the detector.
    pub fn detect(&self, img: &image::GrayImage) -> Vec<Point> {
        detect(img, self)
    }
}

/// One CUDA context per thread: the reference function is re-entrant, the context is not shared.
struct Context(*mut ffi::fdf_ctx);

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::fdf_destroy(self.0) }
    }
}

thread_local! {
    static CONTEXT: RefCell<Option<Context>> = RefCell::new(None);
}

fn with_context<R>(f: impl FnOnce(*mut ffi::fdf_ctx) -> R) -> R {
    CONTEXT.with(|slot| {
        let mut slot = slot.borrow_mut();
        if slot.is_none() {
            let mut ctx = std::ptr::null_mut();
            let st = unsafe { ffi::fdf_create(0, &mut ctx) };
            assert!(st == ffi::FDF_OK, "fdf_create failed: {}", status_string(st));
            *slot = Some(Context(ctx));
        }
        f(slot.as_ref().unwrap().0)
    })
}

fn status_string(st: i32) -> String {
    unsafe { CStr::from_ptr(ffi::fdf_status_string(st)) }.to_string_lossy().into_owned()
}

/// Function to perform the FAST keypoint detection.
pub fn detect(img: &image::GrayImage, config: &Config) -> Vec<Point> {
    // the reference panics here (fast_simd.rs:302-305, :797-801); the C ABI never unwinds, so panic on this side
    assert!(config.count >= 9, "number of consecutive pixels needs to exceed 9");
    assert!(config.count <= 16, "index out of bounds: consecutive count above 16");
    let (w, h) = (img.width(), img.height());
    let worst = (w.saturating_sub(6) as usize) * (h.saturating_sub(6) as usize);
    let nms = match config.non_maximal_supression {
        NonMaximalSuppression::Off => 0u8,
        NonMaximalSuppression::MaxThreshold => 1u8,
        NonMaximalSuppression::SumAbsolute => 2u8,
    };
    let mut out: Vec<Point> = Vec::with_capacity(worst);
    let mut n: usize = 0;
    let st = with_context(|ctx| unsafe {
        ffi::fdf_detect(
            ctx,
            img.as_raw().as_ptr(),
            w,
            h,
            w,
            config.threshold,
            config.count,
            nms,
            out.as_mut_ptr(),
            worst,
            &mut n,
        )
    });
    assert!(st == ffi::FDF_OK, "fdf_detect failed: {}", status_string(st));
    unsafe { out.set_len(n) };
    out.shrink_to_fit();
    out
}
