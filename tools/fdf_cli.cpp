// fdf_cli -- command-line caller of the detector, the counterpart of the reference's src/main.rs:17-83
// built on the C++ host binding (include/fdf.hpp).
//
//   fdf_cli <input.pgm> [output.ppm (default /tmp/output.ppm)] [threshold (16)] [count (9)]
//           [non_maximal_suppression: off|sum_absolute|max_threshold (default sum_absolute, as main.rs:43)]
//
// Same positional arguments, defaults and outputs as main.rs, except that images are binary PGM (P5) in and
// binary PPM (P6) out (the reference decodes/encodes PNG through the `image` crate): the output image is
// the grey input with one pure-red pixel per keypoint (util::draw_plus_sized(.., RED, 1), main.rs:74-77)
// and `<output>.txt` (".ppm" replaced by ".txt") holds "x y\n" per keypoint (main.rs:4-15).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "fdf.hpp"

namespace fd = feature_detector_fast;

static bool read_pgm(const std::string &path, fd::GrayImage &img) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::string magic;
    f >> magic;
    if (magic != "P5") return false;
    auto next_int = [&](long &v) {
        f >> std::ws;
        while (f.peek() == '#') {
            std::string line;
            std::getline(f, line);
            f >> std::ws;
        }
        return (bool)(f >> v);
    };
    long w, h, maxv;
    if (!next_int(w) || !next_int(h) || !next_int(maxv) || maxv != 255 || w <= 0 || h <= 0) return false;
    f.get();  // the single whitespace byte after maxval
    img = fd::GrayImage((uint32_t)w, (uint32_t)h);
    f.read(reinterpret_cast<char *>(img.data.data()), (std::streamsize)img.data.size());
    return (size_t)f.gcount() == img.data.size();
}

int main(int argc, char **argv) {
    if (argc == 1 || (argc == 2 && std::string(argv[1]) == "--help")) {
        std::cout << "fdf_cli <input.pgm> [output(default; /tmp/output.ppm)] [threshold(default: 16)] [count(default:9)] "
                     "[non_maximal_suppression:off|sum_absolute|max_threshold (default: sum_absolute)]\n"
                     " arguments required left to right.\n";
        return 0;
    }
    const std::string input = argv[1];
    const std::string output = argc > 2 ? argv[2] : "/tmp/output.ppm";
    std::string txt = output;
    const size_t pos = txt.rfind(".ppm");
    if (pos != std::string::npos) txt.replace(pos, 4, ".txt");
    else txt += ".txt";
    const int threshold = argc > 3 ? std::atoi(argv[3]) : 16;
    const int count = argc > 4 ? std::atoi(argv[4]) : 9;
    const std::string nms_name = argc > 5 ? argv[5] : "sum_absolute";
    fd::NonMaximalSuppression nms;
    if (nms_name == "off") nms = fd::NonMaximalSuppression::Off;
    else if (nms_name == "sum_absolute") nms = fd::NonMaximalSuppression::SumAbsolute;
    else if (nms_name == "max_threshold") nms = fd::NonMaximalSuppression::MaxThreshold;
    else {
        std::cerr << "unknown non maximal, support: off, sum_absolute, max_threshold\n";
        return 101;
    }
    if (threshold < 0 || threshold > 255 || count < 0 || count > 255) {
        std::cerr << "failed to parse threshold / count\n";
        return 101;
    }
    fd::GrayImage img;
    if (!read_pgm(input, img)) {
        std::cerr << "could not load image at " << input << "\n";
        return 101;
    }
    const fd::Config config{(uint8_t)threshold, (uint8_t)count, nms};
    std::vector<fd::Point> keypoints;
    try {
        const auto t0 = std::chrono::steady_clock::now();
        keypoints = fd::detect(img, config);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "Took: " << ms << "ms, found " << keypoints.size() << " keypoints\n";
    } catch (const std::logic_error &e) {  // where the reference panics
        std::cerr << "panicked: " << e.what() << "\n";
        return 101;
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    std::vector<uint8_t> rgb(img.data.size() * 3);
    for (size_t i = 0; i < img.data.size(); i++) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = img.data[i];
    for (const fd::Point &p : keypoints) {
        const size_t i = ((size_t)p.y * img.width() + p.x) * 3;
        rgb[i] = 255;
        rgb[i + 1] = 0;
        rgb[i + 2] = 0;
    }
    std::ofstream o(output, std::ios::binary);
    o << "P6\n" << img.width() << " " << img.height() << "\n255\n";
    o.write(reinterpret_cast<const char *>(rgb.data()), (std::streamsize)rgb.size());
    std::ofstream t(txt);
    for (const fd::Point &p : keypoints) t << p.x << " " << p.y << "\n";
    return (o && t) ? 0 : 1;
}
