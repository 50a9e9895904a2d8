// facade_demo.cpp -- the reference's "load a .cpb, detect on an image" flow (src/app/acf/acf.cpp:276-360,
// src/test/test-acf-api.cpp:480-506) written against include/acf/ACF.h.
//   facade_demo model.cpb frame.rgb rows cols [nms]
// prints "n" then one "x y w h score" line per detection.  Exit code 2 when the detector is not good().
#include <acf/ACF.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>

int main(int argc, char** argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s model.cpb frame.rgb rows cols [nms]\n", argv[0]); return 1; }
    const int rows = std::atoi(argv[3]), cols = std::atoi(argv[4]);
    acf::Detector detector(argv[1], 0, rows, cols, 1);
    if (!detector.good()) { std::fprintf(stderr, "detector not good: %s\n", detector.error().c_str()); return 2; }
    std::ifstream f(argv[2], std::ios::binary);
    std::vector<unsigned char> px((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if ((int)px.size() != rows * cols * 3) { std::fprintf(stderr, "bad frame file\n"); return 1; }
    acfb_set_hit_capacity(detector.engine(), 1 << 17); // raw cascade hits per frame (default 4096)
    if (argc > 5) { detector.setDoNonMaximaSuppression(true); detector.setMaxDetectionCount(20); }
    ACF_CV::Mat I(rows, cols, 3, 0, px.data());
    std::vector<ACF_CV::Rect> objects;
    std::vector<double> scores;
    try { detector(I, objects, &scores); }
    catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 3; }
    // the reference's other entry points must give the same boxes: a transposed image (setIsTranspose, ACF.h:569-576),
    // a CV_32FC3 image (ACF.cpp:137) and the planar float MatP overload (ACF.h:423-427)
    try
    {
        auto same = [&](const std::vector<ACF_CV::Rect>& o, const std::vector<double>& s, const char* what) {
            bool ok = o.size() == objects.size();
            for (size_t i = 0; ok && i < o.size(); i++)
                ok = o[i].x == objects[i].x && o[i].y == objects[i].y && o[i].width == objects[i].width && o[i].height == objects[i].height && s[i] == scores[i];
            if (!ok) { std::fprintf(stderr, "entry point '%s' disagrees\n", what); std::exit(4); }
        };
        std::vector<unsigned char> pxT((size_t)rows * cols * 3);
        std::vector<float> pxF((size_t)rows * cols * 3);
        acf::MatP planar(cols, rows, 3);
        const float k255 = (float)(1.0 / 255.0);
        for (int y = 0; y < rows; y++)
            for (int x = 0; x < cols; x++)
                for (int c = 0; c < 3; c++)
                {
                    const unsigned char v = px[((size_t)y * cols + x) * 3 + c];
                    pxT[((size_t)x * rows + y) * 3 + c] = v;
                    pxF[((size_t)y * cols + x) * 3 + c] = (float)v * k255;
                    planar.ptr(c)[(size_t)x * rows + y] = (float)v * k255;
                }
        std::vector<ACF_CV::Rect> o2; std::vector<double> s2;
        detector.setIsTranspose(true);
        detector(ACF_CV::Mat(cols, rows, 3, 0, pxT.data()), o2, &s2); same(o2, s2, "transposed");
        detector.setIsTranspose(false);
        o2.clear(); s2.clear();
        detector(ACF_CV::Mat(rows, cols, 3, 5, pxF.data()), o2, &s2); same(o2, s2, "CV_32FC3");
        o2.clear(); s2.clear();
        detector(planar, o2, &s2); same(o2, s2, "MatP");
    }
    catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 3; }
    acf::Detector::Pyramid P;
    detector.computePyramid(I, P);
    {   // the Pyramid overload on the planes just read back, and on the resident copy, must both repeat the boxes
        std::vector<ACF_CV::Rect> o3, o4; std::vector<double> s3, s4;
        detector(P, o3, &s3);
        detector.detectResident(o4, &s4);
        if (o3.size() != objects.size() || o4.size() != objects.size()) { std::fprintf(stderr, "Pyramid overload disagrees\n"); return 4; }
        for (size_t i = 0; i < o3.size(); i++)
            if (o3[i].x != objects[i].x || o3[i].y != objects[i].y || s3[i] != scores[i] || s4[i] != scores[i]) { std::fprintf(stderr, "Pyramid overload disagrees\n"); return 4; }
    }
    try
    {   // the stand-alone operators (Detector::rgbConvert / convTri / gradientMag / gradientHist, ACF.h:443-491) chained the way
        // chnsCompute chains them (chnsCompute.cpp:228-338) must rebuild the histogram channels of the resident real scale 0
        const acfb_options& o = detector.options();
        if (o.color_space == 0 && rows % 4 == 0 && cols % 4 == 0)
        {
            acf::MatP rgb(cols, rows, 3), g, M, O, H;
            const float k255 = (float)(1.0 / 255.0);
            for (int y = 0; y < rows; y++)
                for (int x = 0; x < cols; x++)
                    for (int c = 0; c < 3; c++) rgb.ptr(c)[(size_t)x * rows + y] = (float)px[((size_t)y * cols + x) * 3 + c] * k255;
            detector.rgbConvert(rgb, g, "gray");
            detector.convTri(g, g, o.color_smooth);
            detector.gradientMag(g, M, O, o.gm_colorChn, o.gm_normRad, o.gm_normConst, o.gm_full);
            detector.gradientHist(M, O, H, o.shrink, o.gh_nOrients, o.gh_softBin, o.gh_useHog, o.gh_clipHog, o.gm_full);
            const int ch = rows / 4, cw = cols / 4, nch = 1 + o.gh_nOrients + (o.color_enabled ? 1 : 0);
            std::vector<float> R((size_t)nch * ch * cw);
            int d = 0, w = 0, h = 0;
            detector.computePyramid(I, P); // makes this frame's real-scale channels resident again
            if (acfb_tap(detector.engine(), "R", 0, 0, R.data(), R.size(), &d, &w, &h) != 0) { std::fprintf(stderr, "tap: %s\n", acfb_last_error()); return 5; }
            const float* Rh = R.data() + (size_t)(d - o.gh_nOrients) * w * h;
            if (w != H.rows() || h != H.cols() || std::memcmp(Rh, H.ptr(), (size_t)o.gh_nOrients * w * h * sizeof(float)) != 0)
            { std::fprintf(stderr, "operator chain disagrees with the resident channels\n"); return 5; }
        }
    }
    catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 3; }
    std::printf("%zu %d\n", objects.size(), P.nScales);
    for (size_t i = 0; i < objects.size(); i++)
        std::printf("%d %d %d %d %.9g\n", objects[i].x, objects[i].y, objects[i].width, objects[i].height, scores[i]);
    return 0;
}
