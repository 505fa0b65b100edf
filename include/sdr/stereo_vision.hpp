// C++17 facade over the C ABI (sdr_c.h) with the reference's class and method names:
//   StereoConfiguration  stereo_vision/include/stereo_configuration.hpp:6-16
//   StereoRectifier      stereo_vision/include/stereo_rectifier.hpp:7-19   (ctor + rectify)
//   StereoDisparity      stereo_vision/include/stereo_disparity.hpp:8-24   (computeDisparity, show_disparityMap, computeDepth,
//                        show_depthMap, get_matcher; the WLS filter is NOT provided -- see the class comment)
//   convertCVMatToPCL + pcl::VoxelGrid usage of point_cloud/src/pcd_write.cpp:17-51, :123-130 -> VoxelGrid / PointCloud
// OpenCV is not a dependency: images travel as sdr::Mat, whose {data, rows, cols, step} fields are
// layout-compatible with cv::Mat's, so `sdr::Mat::view(m.data, m.rows, m.cols, m.channels(), m.elemSize1(), m.step)`
// adapts a cv::Mat in one line where OpenCV exists.
// Constructors come in the reference's shapes (StereoRectifier(config), StereoDisparity(Q), StereoSGBM::create(...),
// VoxelGrid()) on a process-wide default context, plus overloads that take an explicit sdr::Context (one per device).
// Error behaviour mirrors the reference: loadFromFile/saveToFile return bool and print to stderr; bad arguments
// throw sdr::Error (where OpenCV would throw cv::Exception); rectify() on empty input prints and returns.
#pragma once
#include <array>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../sdr_c.h"

namespace sdr {

class Error : public std::runtime_error {
  public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }

  private:
    int status_;
};

class Context {
  public:
    explicit Context(int device = 0) {
        int rc = sdr_create(device, &h_);
        if (rc != SDR_OK) throw Error(rc, "sdr_create failed: no usable sm_100 CUDA device (there is no CPU fallback)");
    }
    ~Context() { sdr_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sdr_ctx* get() const { return h_; }
    // The context behind the reference-shaped constructors (StereoRectifier(config), StereoDisparity(Q), ...): device 0,
    // created on first use, alive until the process exits -- the reference has no notion of a device either.
    static Context& shared() {
        static Context ctx(0);
        return ctx;
    }
    void check(int rc, bool allow_voxel_overflow = false) const {
        if (rc == SDR_OK || (allow_voxel_overflow && rc == SDR_EVOXEL_OVERFLOW)) return;
        throw Error(rc, sdr_last_error(h_));
    }

  private:
    sdr_ctx* h_ = nullptr;
};

struct Size {
    int width = 0, height = 0;
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};

// Minimal image container / view (host memory, or a non-owning view of host or device memory).
class Mat {
  public:
    Mat() = default;
    Mat(int rows, int cols, int channels, int elem_size) { create(rows, cols, channels, elem_size); }
    static Mat view(void* data, int rows, int cols, int channels, int elem_size, size_t step, int mem = SDR_MEM_HOST) {
        Mat m;
        m.im_ = sdr_image{data, rows, cols, channels, elem_size, step, mem};
        return m;
    }
    void create(int rows, int cols, int channels, int elem_size) {
        if (!own_ || im_.rows != rows || im_.cols != cols || im_.channels != channels || im_.elem_size != elem_size) {
            own_ = std::make_shared<std::vector<unsigned char>>((size_t)rows * cols * channels * elem_size);
            im_ = sdr_image{own_->data(), rows, cols, channels, elem_size, (size_t)cols * channels * elem_size, SDR_MEM_HOST};
        }
    }
    bool empty() const { return im_.data == nullptr || im_.rows <= 0 || im_.cols <= 0; }
    Size size() const { return Size{im_.cols, im_.rows}; }
    int rows() const { return im_.rows; }
    int cols() const { return im_.cols; }
    int channels() const { return im_.channels; }
    template <class T> T* ptr(int y = 0) { return (T*)((unsigned char*)im_.data + (size_t)y * im_.step_bytes); }
    template <class T> const T* ptr(int y = 0) const { return (const T*)((const unsigned char*)im_.data + (size_t)y * im_.step_bytes); }
    // ROI view, as cv::Mat::operator()(cv::Rect) at stereo_displayer.cpp:155-156
    Mat roi(int x, int y, int w, int h) const {
        Mat m = *this;
        m.im_.data = (unsigned char*)im_.data + (size_t)y * im_.step_bytes + (size_t)x * im_.channels * im_.elem_size;
        m.im_.cols = w;
        m.im_.rows = h;
        return m;
    }
    const sdr_image* c() const { return &im_; }
    sdr_image* c() { return &im_; }

  private:
    sdr_image im_{};
    std::shared_ptr<std::vector<unsigned char>> own_;
};

struct StereoConfiguration {
    std::array<double, 9> cameraMatrixLeft{}, cameraMatrixRight{}, R{}, E{}, F{}, R1{}, R2{};
    std::array<double, 5> distCoeffsLeft{}, distCoeffsRight{};
    std::array<double, 3> T{};
    std::array<double, 12> P1{}, P2{};
    std::array<double, 16> Q{};
    Size imageSize;
    unsigned present = 0;

    bool loadFromFile(const std::string& filename) {
        sdr_calibration c;
        if (sdr_load_calibration_yaml(filename.c_str(), &c) != SDR_OK) {
            std::cerr << "Error: Failed to load stereo configuration from " << filename << std::endl;
            return false;
        }
        from_c(c);
        std::cout << "Stereo configuration loaded successfully from " << filename << std::endl;
        return true;
    }
    bool saveToFile(const std::string& filename) const {
        sdr_calibration c = to_c();
        return sdr_save_calibration_yaml(filename.c_str(), &c) == SDR_OK;
    }
    bool isValid() const {
        sdr_calibration c = to_c();
        return sdr_calibration_is_valid(&c) != 0;
    }
    sdr_calibration to_c() const {
        sdr_calibration c{};
        c.image_width = imageSize.width;
        c.image_height = imageSize.height;
        std::memcpy(c.K_left, cameraMatrixLeft.data(), sizeof c.K_left);
        std::memcpy(c.dist_left, distCoeffsLeft.data(), sizeof c.dist_left);
        std::memcpy(c.K_right, cameraMatrixRight.data(), sizeof c.K_right);
        std::memcpy(c.dist_right, distCoeffsRight.data(), sizeof c.dist_right);
        std::memcpy(c.R, R.data(), sizeof c.R);
        std::memcpy(c.T, T.data(), sizeof c.T);
        std::memcpy(c.E, E.data(), sizeof c.E);
        std::memcpy(c.F, F.data(), sizeof c.F);
        std::memcpy(c.R1, R1.data(), sizeof c.R1);
        std::memcpy(c.R2, R2.data(), sizeof c.R2);
        std::memcpy(c.P1, P1.data(), sizeof c.P1);
        std::memcpy(c.P2, P2.data(), sizeof c.P2);
        std::memcpy(c.Q, Q.data(), sizeof c.Q);
        c.present = present;
        return c;
    }
    void from_c(const sdr_calibration& c) {
        imageSize = Size{c.image_width, c.image_height};
        std::memcpy(cameraMatrixLeft.data(), c.K_left, sizeof c.K_left);
        std::memcpy(distCoeffsLeft.data(), c.dist_left, sizeof c.dist_left);
        std::memcpy(cameraMatrixRight.data(), c.K_right, sizeof c.K_right);
        std::memcpy(distCoeffsRight.data(), c.dist_right, sizeof c.dist_right);
        std::memcpy(R.data(), c.R, sizeof c.R);
        std::memcpy(T.data(), c.T, sizeof c.T);
        std::memcpy(E.data(), c.E, sizeof c.E);
        std::memcpy(F.data(), c.F, sizeof c.F);
        std::memcpy(R1.data(), c.R1, sizeof c.R1);
        std::memcpy(R2.data(), c.R2, sizeof c.R2);
        std::memcpy(P1.data(), c.P1, sizeof c.P1);
        std::memcpy(P2.data(), c.P2, sizeof c.P2);
        std::memcpy(Q.data(), c.Q, sizeof c.Q);
        present = c.present;
    }
};

class StereoRectifier {
  public:
    // stereo_rectifier.hpp:15: explicit StereoRectifier(const StereoConfiguration &config)
    explicit StereoRectifier(const StereoConfiguration& config) : StereoRectifier(config, Context::shared()) {}
    StereoRectifier(const StereoConfiguration& config, Context& ctx) : config_(config), ctx_(ctx) {
        sdr_calibration c = config.to_c();
        ctx_.check(sdr_rectifier_create(ctx_.get(), &c, &h_));
        std::cout << "Remapping tables created." << std::endl;  // evaluated inside the kernel; kept for log parity
    }
    ~StereoRectifier() { sdr_rectifier_destroy(h_); }
    StereoRectifier(const StereoRectifier&) = delete;
    void rectify(const Mat& left_src, const Mat& right_src, Mat& left_dst, Mat& right_dst) {
        if (left_src.empty() || right_src.empty()) {
            std::cerr << "Error: Input images are empty" << std::endl;
            return;
        }
        left_dst.create(config_.imageSize.height, config_.imageSize.width, left_src.channels(), 1);
        right_dst.create(config_.imageSize.height, config_.imageSize.width, right_src.channels(), 1);
        ctx_.check(sdr_rectify(h_, left_src.c(), right_src.c(), left_dst.c(), right_dst.c(), nullptr, nullptr));
    }

  private:
    const StereoConfiguration& config_;
    Context& ctx_;
    sdr_rectifier* h_ = nullptr;
};

// cv::StereoSGBM subset the reference uses
class StereoSGBM {
  public:
    enum { MODE_SGBM = 0, MODE_HH = 1, MODE_SGBM_3WAY = 2 };
    static std::shared_ptr<StereoSGBM> create(Context& ctx, int minDisparity = 0, int numDisparities = 16, int blockSize = 3, int P1 = 0,
                                              int P2 = 0, int disp12MaxDiff = 0, int preFilterCap = 0, int uniquenessRatio = 0,
                                              int speckleWindowSize = 0, int speckleRange = 0, int mode = MODE_SGBM) {
        sdr_sgbm_params p{minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap, uniquenessRatio,
                          speckleWindowSize, speckleRange, mode};
        return std::shared_ptr<StereoSGBM>(new StereoSGBM(ctx, p));
    }
    // cv::StereoSGBM::create(minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap, uniquenessRatio,
    // speckleWindowSize, speckleRange, mode): same defaults, the shared context
    static std::shared_ptr<StereoSGBM> create(int minDisparity = 0, int numDisparities = 16, int blockSize = 3, int P1 = 0, int P2 = 0,
                                              int disp12MaxDiff = 0, int preFilterCap = 0, int uniquenessRatio = 0, int speckleWindowSize = 0,
                                              int speckleRange = 0, int mode = MODE_SGBM) {
        return create(Context::shared(), minDisparity, numDisparities, blockSize, P1, P2, disp12MaxDiff, preFilterCap, uniquenessRatio,
                      speckleWindowSize, speckleRange, mode);
    }
    // cv::ximgproc::createRightMatcher(matcher_left) (stereo_disparity.cpp:10)
    static std::shared_ptr<StereoSGBM> createRightMatcher(const std::shared_ptr<StereoSGBM>& left) {
        sdr_sgbm_params r;
        sdr_sgbm_right_matcher_params(&left->p_, &r);
        return std::shared_ptr<StereoSGBM>(new StereoSGBM(left->ctx_, r));
    }
    ~StereoSGBM() { sdr_sgbm_destroy(h_); }
    void compute(const Mat& left, const Mat& right, Mat& disp) {
        disp.create(left.rows(), left.cols(), 1, 2);
        ctx_.check(sdr_sgbm_compute(h_, left.c(), right.c(), disp.c()));
    }
    int getNumDisparities() const { return p_.numDisparities; }
    int getMinDisparity() const { return p_.minDisparity; }

  private:
    StereoSGBM(Context& ctx, const sdr_sgbm_params& p) : ctx_(ctx), p_(p) { ctx_.check(sdr_sgbm_create(ctx_.get(), &p, &h_)); }
    Context& ctx_;
    sdr_sgbm_params p_;
    sdr_sgbm* h_ = nullptr;
};

class StereoDisparity {
  public:
    // stereo_disparity.hpp:16 / stereo_disparity.cpp:4-14: StereoDisparity(const cv::Mat &Q_matrix) -- the reference's
    // parameter tuple (0, 80, 5, 600, 2400, 1, 63, 12, 200, 2, MODE_SGBM_3WAY) and its right matcher.
    // NOT provided: cv::ximgproc::DisparityWLSFilter (:11-13, :31). ximgproc exists in no oracle available to this build,
    // so a restatement could not be pinned; computeDisparity therefore returns the matcher's UNFILTERED left disparity
    // where the reference returns the WLS-filtered one (same size, same type, different values). wlsFilterApplied() says so.
    explicit StereoDisparity(const std::array<double, 16>& Q_matrix) : StereoDisparity(Q_matrix, Context::shared()) {}
    StereoDisparity(const std::array<double, 16>& Q_matrix, Context& ctx) : ctx_(ctx), Q_(Q_matrix) {
        matcher_ = StereoSGBM::create(ctx, 0, 80, 5, 8 * 5 * 5 * 3, 32 * 5 * 5 * 3, 1, 63, 12, 200, 2, StereoSGBM::MODE_SGBM_3WAY);
        right_matcher_ = StereoSGBM::createRightMatcher(matcher_);
        ctx_.check(sdr_vis_create(ctx_.get(), &vis_));
    }
    ~StereoDisparity() { sdr_vis_destroy(vis_); }
    StereoDisparity(const StereoDisparity&) = delete;
    static constexpr bool wlsFilterApplied() { return false; }
    // stereo_disparity.cpp:17-39 without the WLS filter: BGR2GRAY, half-size INTER_AREA, matcher on (left, right) and the
    // right matcher on (right, left) (kept in lastRightDisparity(), what wls_filter->filter would consume), convertTo
    // (CV_32F, 1/16). Returns the HALF-resolution CV_32F map, like the reference. left / right: BGR or gray, even size.
    Mat computeDisparity(const Mat& left, const Mat& right) {
        Mat lg = gray(left), rg = gray(right);
        Mat ls(lg.rows() / 2, lg.cols() / 2, 1, 1), rs(rg.rows() / 2, rg.cols() / 2, 1, 1);
        ctx_.check(sdr_resize_half_area(ctx_.get(), lg.c(), ls.c()));
        ctx_.check(sdr_resize_half_area(ctx_.get(), rg.c(), rs.c()));
        Mat disp_left;
        matcher_->compute(ls, rs, disp_left);
        right_matcher_->compute(rs, ls, disp_right_);
        Mat out(disp_left.rows(), disp_left.cols(), 1, 4);
        ctx_.check(sdr_disp_to_float(ctx_.get(), disp_left.c(), out.c()));
        return out;
    }
    const Mat& lastRightDisparity() const { return disp_right_; }   // CV_16S, as right_matcher->compute gives it
    // The un-resized path of point_cloud/src/pcd_write.cpp:88-112 (gray, matcher, convertTo 1/16) at full resolution.
    Mat computeDisparityFullRes(const Mat& left, const Mat& right) {
        Mat lg = gray(left), rg = gray(right), d16;
        matcher_->compute(lg, rg, d16);
        Mat out(d16.rows(), d16.cols(), 1, 4);
        ctx_.check(sdr_disp_to_float(ctx_.get(), d16.c(), out.c()));   // convertTo(CV_32F, 1/16) on the device
        return out;
    }
    // stereo_disparity.cpp:42-73
    Mat show_disparityMap(const Mat& disparity) {
        Mat out(disparity.rows(), disparity.cols(), 1, 1);
        ctx_.check(sdr_vis_disparity(vis_, disparity.c(), matcher_->getNumDisparities(), out.c()));
        return out;
    }
    // cv::reprojectImageTo3D(disparity, depth, Q)  (stereo_disparity.cpp:76-80)
    Mat computeDepth(const Mat& disparity, bool handleMissingValues = false) {
        Mat depth(disparity.rows(), disparity.cols(), 3, 4);
        ctx_.check(sdr_reproject_to_3d(ctx_.get(), disparity.c(), Q_.data(), handleMissingValues ? 1 : 0, depth.c()));
        return depth;
    }
    // stereo_disparity.cpp:83-124 (the argument is the depth image, 3-channel points or 1-channel Z)
    Mat show_depthMap(const Mat& depth) {
        Mat out(depth.rows(), depth.cols(), 3, 1);
        ctx_.check(sdr_vis_depth(vis_, depth.c(), out.c()));
        return out;
    }
    const std::shared_ptr<StereoSGBM> get_matcher() const { return matcher_; }

  private:
    Mat gray(const Mat& m) {
        if (m.channels() == 1) return m;
        Mat g(m.rows(), m.cols(), 1, 1);
        ctx_.check(sdr_bgr2gray(ctx_.get(), m.c(), g.c()));
        return g;
    }
    Context& ctx_;
    std::array<double, 16> Q_;
    std::shared_ptr<StereoSGBM> matcher_, right_matcher_;
    Mat disp_right_;
    sdr_vis* vis_ = nullptr;
};

// pcl::PointCloud<pcl::PointXYZRGB> stand-in: what convertCVMatToPCL + VoxelGrid::filter produce.
struct PointCloud {
    std::vector<float> xyz;          // n x 3
    std::vector<unsigned char> rgb;  // n x 3 (r, g, b)
    size_t width = 0, height = 1;
    bool is_dense = false;
    size_t size() const { return xyz.size() / 3; }
};

class VoxelGrid {
  public:
    VoxelGrid() : VoxelGrid(Context::shared()) {}   // pcl::VoxelGrid<pcl::PointXYZRGB> sor;  (pcd_write.cpp:123)
    explicit VoxelGrid(Context& ctx) : ctx_(ctx) { ctx_.check(sdr_cloud_create(ctx_.get(), &h_)); }
    ~VoxelGrid() { sdr_cloud_destroy(h_); }
    void setLeafSize(float lx, float ly, float lz) { opt_.leaf[0] = lx; opt_.leaf[1] = ly; opt_.leaf[2] = lz; }
    void setDropInvalid(bool on, int min_disparity = 0) { opt_.drop_invalid = on; opt_.min_disparity = min_disparity; }
    void setZMax(float z) { opt_.z_max = z; }
    // convertCVMatToPCL(pointCloud_CV, colorImage_CV) followed by filter(). Returns false when PCL's int32 index guard
    // trips ("Leaf size is too small"): the cloud is then the unfiltered organised input, as PCL returns it.
    bool filter(const Mat& pointCloud_CV, const Mat& colorImage_CV, PointCloud& out, const Mat* disp16 = nullptr) {
        sdr_cloud_view v{};
        int rc = sdr_voxelgrid(h_, pointCloud_CV.c(), colorImage_CV.empty() ? nullptr : colorImage_CV.c(), disp16 ? disp16->c() : nullptr,
                               &opt_, &v);
        ctx_.check(rc, true);
        out.xyz.resize((size_t)v.n * 3);
        out.rgb.resize((size_t)v.n * 3);
        ctx_.check(sdr_cloud_download(h_, out.xyz.data(), out.rgb.data(), nullptr, nullptr));
        out.width = v.passthrough ? (size_t)pointCloud_CV.cols() : (size_t)v.n;
        out.height = v.passthrough ? (size_t)pointCloud_CV.rows() : 1;
        out.is_dense = !v.passthrough;
        if (rc == SDR_EVOXEL_OVERFLOW) std::cerr << "[pcl::VoxelGrid::applyFilter] " << sdr_last_error(ctx_.get()) << std::endl;
        return rc == SDR_OK;
    }

  private:
    Context& ctx_;
    sdr_cloud* h_ = nullptr;
    sdr_voxel_options opt_{{0.005f, 0.005f, 0.005f}, 0, 0.f, 0};
};

// Frame ingest in front of the loop of stereo_displayer.cpp:145-156 (`cap >> frame`): a decoder surface (NV12) or a UVC camera
// buffer (YUYV) becomes the BGR side-by-side frame cv::VideoCapture would have returned: cv::cvtColor(src, bgr,
// COLOR_YUV2BGR_NV12 / COLOR_YUV2BGR_YUY2), bit-exact, on the device. format: SDR_PIX_NV12 or SDR_PIX_YUYV.
inline void cvtColorToBGR(Context& ctx, const Mat& src, Mat& bgr, int format) {
    const int rows = format == SDR_PIX_NV12 ? src.rows() * 2 / 3 : src.rows();
    bgr.create(rows, src.cols(), 3, 1);
    ctx.check(sdr_ingest_to_bgr(ctx.get(), format, src.c(), bgr.c()));
}
inline void cvtColorToBGR(const Mat& src, Mat& bgr, int format) { cvtColorToBGR(Context::shared(), src, bgr, format); }

}  // namespace sdr
