// fdf.hpp -- C++ host-side mirror of the reference crate's public API for the detection path
// (reference: src/lib.rs:15-64), implemented over the C ABI of include/fdf.h.
//
// The reference is a Rust crate and there is no Rust toolchain in the build image, so this header is
// the compiled-language host binding that is actually built and tested; rust/ holds the equivalent
// (uncompiled here) Rust crate.  Same names, same argument meaning, same error behaviour:
//   - `Config{threshold, count, non_maximal_supression}.detect(img)` and `detect(img, config)`
//     return the keypoints in the reference's row-major order;
//   - count outside 9..=16 "panics" (reference: assert at fast_simd.rs:302-305, index panic at :797-801)
//     -> std::logic_error; CUDA / device failures -> std::runtime_error.  There is no CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "fdf.h"

namespace feature_detector_fast {

// lib.rs:15-20
struct Point {
    uint32_t x = 0;
    uint32_t y = 0;
    bool operator==(const Point &o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point &o) const { return !(*this == o); }
};
static_assert(sizeof(Point) == sizeof(fdf_point), "Point must be layout-compatible with fdf_point");

// lib.rs:25-36 (values as fast_simd.rs:74-76)
enum class NonMaximalSuppression : uint8_t { Off = 0, MaxThreshold = 1, SumAbsolute = 2 };

// The part of image::GrayImage the detection path uses: contiguous row-major u8, pitch == width
// (fast_simd.rs:307-308, 330).
struct GrayImage {
    uint32_t width_ = 0, height_ = 0;
    std::vector<uint8_t> data;
    GrayImage() = default;
    GrayImage(uint32_t w, uint32_t h) : width_(w), height_(h), data((size_t)w * h, 0) {}
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    const std::vector<uint8_t> &as_raw() const { return data; }
    uint8_t &at(uint32_t x, uint32_t y) { return data[(size_t)y * width_ + x]; }
};

namespace detail {
struct Context {
    fdf_ctx *ctx = nullptr;
    std::vector<Point> scratch;
    Context() {
        fdf_status st = fdf_create(0, &ctx);
        if (st != FDF_OK) throw std::runtime_error(std::string("fdf_create: ") + fdf_status_string(st));
    }
    ~Context() { fdf_destroy(ctx); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
};
inline Context &thread_context() {  // the reference function is re-entrant; the context is per thread
    thread_local Context c;
    return c;
}
}  // namespace detail

struct Config;
std::vector<Point> detect(const GrayImage &img, const Config &config);

// lib.rs:38-59
struct Config {
    uint8_t threshold;
    uint8_t count;
    NonMaximalSuppression non_maximal_supression;  // spelled as in the reference
    std::vector<Point> detect(const GrayImage &img) const { return feature_detector_fast::detect(img, *this); }
};

// lib.rs:62-64
inline std::vector<Point> detect(const GrayImage &img, const Config &config) {
    if (config.count < 9) throw std::logic_error("number of consecutive pixels needs to exceed 9");
    if (config.count > 16) throw std::logic_error("index out of bounds: consecutive count above 16");
    detail::Context &c = detail::thread_context();
    const size_t w = img.width(), h = img.height();
    const size_t worst = (w > 6 && h > 6) ? (w - 6) * (h - 6) : 0;
    if (c.scratch.size() < worst + 1) c.scratch.resize(worst + 1);
    size_t n = 0;
    fdf_status st = fdf_detect(c.ctx, img.as_raw().data(), img.width(), img.height(), img.width(), config.threshold,
                               config.count, (uint8_t)config.non_maximal_supression,
                               reinterpret_cast<fdf_point *>(c.scratch.data()), worst, &n);
    if (st == FDF_ERR_INVALID_COUNT) throw std::logic_error(fdf_last_error(c.ctx));
    if (st != FDF_OK) throw std::runtime_error(std::string("fdf_detect: ") + fdf_last_error(c.ctx));
    return std::vector<Point>(c.scratch.begin(), c.scratch.begin() + (ptrdiff_t)n);
}

// Streaming form (no counterpart in the reference, whose `detect` is synchronous): up to `depth` images in flight on
// this thread's context, results first in, first out (C ABI: fdf_pipe_*, include/fdf.h).
//     Pipe pipe(4, 1920, 1080);
//     for (const GrayImage &img : frames) {
//         if (pipe.in_flight() == pipe.depth()) handle(pipe.collect());
//         pipe.submit(img, config);
//     }
//     while (pipe.in_flight()) handle(pipe.collect());
class Pipe {
public:
    Pipe(uint32_t depth, uint32_t max_width, uint32_t max_height, size_t cap = 0) : depth_(depth) {
        const size_t worst = (max_width > 6 && max_height > 6) ? (size_t)(max_width - 6) * (max_height - 6) : 0;
        cap_ = cap ? cap : (worst / 16 > 4096 ? worst / 16 : (worst < 4096 ? worst : 4096));
        detail::Context &c = detail::thread_context();
        fdf_status st = fdf_pipe_create(c.ctx, depth, max_width, max_height, cap_, &pipe_);
        if (st != FDF_OK) throw std::runtime_error(std::string("fdf_pipe_create: ") + fdf_last_error(c.ctx));
        out_.resize(cap_ + 1);
    }
    ~Pipe() { fdf_pipe_destroy(pipe_); }
    Pipe(const Pipe &) = delete;
    Pipe &operator=(const Pipe &) = delete;
    uint32_t depth() const { return depth_; }
    uint32_t in_flight() const { return fdf_pipe_in_flight(pipe_); }
    void submit(const GrayImage &img, const Config &config) {
        if (config.count < 9) throw std::logic_error("number of consecutive pixels needs to exceed 9");
        if (config.count > 16) throw std::logic_error("index out of bounds: consecutive count above 16");
        fdf_status st = fdf_pipe_submit(pipe_, img.as_raw().data(), img.width(), img.height(), img.width(),
                                        config.threshold, config.count, (uint8_t)config.non_maximal_supression);
        if (st != FDF_OK) throw std::runtime_error(std::string("fdf_pipe_submit: ") + fdf_last_error(detail::thread_context().ctx));
    }
    std::vector<Point> collect() {
        size_t n = 0;
        fdf_status st = fdf_pipe_collect(pipe_, reinterpret_cast<fdf_point *>(out_.data()), cap_, &n);
        if (st != FDF_OK) throw std::runtime_error(std::string("fdf_pipe_collect: ") + fdf_last_error(detail::thread_context().ctx));
        return std::vector<Point>(out_.begin(), out_.begin() + (ptrdiff_t)n);
    }

private:
    fdf_pipe *pipe_ = nullptr;
    uint32_t depth_;
    size_t cap_ = 0;
    std::vector<Point> out_;
};

}  // namespace feature_detector_fast
