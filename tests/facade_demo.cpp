// A maintainer-style program against include/sdr/stereo_vision.hpp, written like the reference's own callers
// (stereo_displayer.cpp:125-162 and pcd_write.cpp:82-130): load the calibration, split an SBS frame by ROI views, rectify,
// computeDisparity (live half-size path) and the full-size pcd_write path, computeDepth, VoxelGrid::filter, the two
// visualisations. Every buffer is dumped so tests/test_gpu_facade.py can compare it with the oracle / cv2.
// usage: facade_demo calib.yaml sbs.raw(2560x720x3 u8) outdir
#include <cstdio>
#include <string>
#include <vector>

#include "sdr/stereo_vision.hpp"

static void dump(const std::string& path, const void* p, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb");
    fwrite(p, 1, bytes, f);
    fclose(f);
}
static void dump(const std::string& path, const sdr::Mat& m, int elem) {
    std::vector<unsigned char> buf((size_t)m.rows() * m.cols() * m.channels() * elem);
    for (int y = 0; y < m.rows(); y++)
        memcpy(&buf[(size_t)y * m.cols() * m.channels() * elem], m.ptr<unsigned char>(y), (size_t)m.cols() * m.channels() * elem);
    dump(path, buf.data(), buf.size());
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string out = argv[3];
    sdr::StereoConfiguration config;
    if (!config.loadFromFile(argv[1])) return 3;
    const int W = config.imageSize.width, H = config.imageSize.height;
    std::vector<unsigned char> frame((size_t)H * 2 * W * 3);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(frame.data(), 1, frame.size(), f) != frame.size()) return 4;
    fclose(f);
    try {
        // the reference's constructor shapes: no context argument anywhere
        sdr::StereoRectifier rectifier(config);
        sdr::StereoDisparity disparity(config.Q);
        sdr::Mat sbs = sdr::Mat::view(frame.data(), H, 2 * W, 3, 1, (size_t)2 * W * 3);
        sdr::Mat left_raw = sbs.roi(0, 0, W, H), right_raw = sbs.roi(W, 0, W, H);   // stereo_displayer.cpp:155-156
        sdr::Mat left_rect, right_rect;
        rectifier.rectify(left_raw, right_raw, left_rect, right_rect);
        dump(out + "/left_rect.u8", left_rect, 1);
        dump(out + "/right_rect.u8", right_rect, 1);
        sdr::Mat disp_half = disparity.computeDisparity(left_rect, right_rect);     // live path, half size, no WLS
        dump(out + "/disp_half.f32", disp_half, 4);
        dump(out + "/disp_right_half.s16", disparity.lastRightDisparity(), 2);
        sdr::Mat vis = disparity.show_disparityMap(disp_half);
        dump(out + "/vis_disp.u8", vis, 1);
        sdr::Mat disp = disparity.computeDisparityFullRes(left_rect, right_rect);   // pcd_write.cpp:88-112
        dump(out + "/disp.f32", disp, 4);
        sdr::Mat depth = disparity.computeDepth(disp, true);                        // pcd_write.cpp:116
        dump(out + "/depth.f32", depth, 4);
        sdr::Mat dvis = disparity.show_depthMap(depth);
        dump(out + "/vis_depth.u8", dvis, 1);
        sdr::VoxelGrid sor;                                                          // pcd_write.cpp:123-130
        sor.setLeafSize(5.0f, 5.0f, 5.0f);
        sor.setZMax(5000.f);
        sdr::PointCloud cloud;
        const bool filtered = sor.filter(depth, left_rect, cloud);
        dump(out + "/cloud_xyz.f32", cloud.xyz.data(), cloud.xyz.size() * 4);
        dump(out + "/cloud_rgb.u8", cloud.rgb.data(), cloud.rgb.size());
        printf("ok filtered=%d points=%zu half=%dx%d matcher_numdisp=%d wls=%d\n", (int)filtered, cloud.size(), disp_half.cols(), disp_half.rows(),
               disparity.get_matcher()->getNumDisparities(), (int)sdr::StereoDisparity::wlsFilterApplied());
    } catch (const sdr::Error& e) {
        fprintf(stderr, "sdr::Error %d: %s\n", e.status(), e.what());
        return 10 + e.status();
    }
    return 0;
}
