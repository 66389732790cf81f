// surf_demo.cpp -- the reference's demo flow (main.cpp:163-283, cudaSurfDemo2) on this library, without
// OpenCV: load or synthesise a stereo pair, upload, 100x {detect+describe left, right}, 100x match, print
// counts and times. Written against include/compat/surf.h, i.e. the reference's own surf.h interface.
//
//   surf_demo [device] [left.pgm right.pgm]        (no files: two 1280x960 synth_v1 frames 12 px apart)
// Environment: SURFB200_FRESH_DESC=1 keeps the reference's descriptor-buffer contract (a new cudaMalloc per call,
// surfd.cu:3262-3266); SURF_DEMO_REPEATS=n overrides the 100 repetitions of main.cpp:239-251. The last lines are
// machine-readable (`key=value`) for tests/test_demo_gpu.py.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "surf.h"

static bool read_pgm(const std::string& path, std::vector<unsigned char>& px, int& w, int& h) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv = 0;
    if (!(f >> magic) || magic != "P5") return false;
    auto skip = [&]() { while (f >> std::ws && f.peek() == '#') f.ignore(1 << 20, '\n'); };
    skip(); f >> w; skip(); f >> h; skip(); f >> maxv;
    f.get();
    if (!f || maxv != 255 || w <= 0 || h <= 0) return false;
    px.resize((size_t)w * h);
    f.read((char*)px.data(), px.size());
    return (bool)f;
}

int main(int argc, char** argv) {
    const int devNum = argc > 1 ? std::atoi(argv[1]) : 0;
    std::vector<unsigned char> limg, rimg;
    int w = 1280, h = 960;
    if (argc > 3) {
        int w2 = 0, h2 = 0;
        if (!read_pgm(argv[2], limg, w, h) || !read_pgm(argv[3], rimg, w2, h2) || w2 != w || h2 != h) {
            std::fprintf(stderr, "cannot read the P5 pair\n");
            return 1;
        }
    } else {
        limg.resize((size_t)w * h); rimg.resize((size_t)w * h);
        sb_synth_frame(limg.data(), w, h, w, 5000, 0, 0, 0);
        sb_synth_frame(rimg.data(), w, h, w, 5000, 12, 2, 5000 ^ 0xA5A5);
    }
    std::cout << "Image size = (" << w << "," << h << ")" << std::endl;

    // main.cpp:187-204
    const int samplingStep = 2, octaves = 4, initLobe = 3, indexSize = 4, max_npts = 10000;
    const float thres = 4.f;
    const bool doubleImageSize = false, upright = true, extended = false;

    std::cout << "Initializing data..." << std::endl;
    initDevice(devNum);
    GpuTimer timer(0);
    int3 whp;
    whp.x = w; whp.y = h; whp.z = iAlignUp(w, 128);
    unsigned char *img1 = NULL, *img2 = NULL;
    CHECK(cudaMalloc((void**)&img1, (size_t)whp.z * h));
    CHECK(cudaMalloc((void**)&img2, (size_t)whp.z * h));
    CHECK(cudaMemcpy2D(img1, whp.z, limg.data(), w, w, h, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy2D(img2, whp.z, rimg.data(), w, w, h, cudaMemcpyHostToDevice));
    const float t0 = timer.read();

    surf::SurfData d1, d2;
    surf::initSurfData(d1, max_npts, true, true);
    surf::initSurfData(d2, max_npts, true, true);
    float *desc1 = NULL, *desc2 = NULL;
    std::unique_ptr<surf::Surfor> detector(new surf::Surfor);
    detector->init(octaves, thres, doubleImageSize, initLobe * 3, samplingStep, upright, extended, indexSize, w, h);

    const int nrepeats = std::getenv("SURF_DEMO_REPEATS") ? std::max(1, std::atoi(std::getenv("SURF_DEMO_REPEATS"))) : 100;
    detector->detectAndCompute(img1, d1, whp, &desc1, true);  // context creation outside the timed loop
    // descriptor-buffer contract: what happens to the first pointer when the same variable is passed again
    float* first_ptr = desc1;
    std::vector<float> first_copy((size_t)d1.num_pts * 64);
    CHECK(cudaMemcpy(first_copy.data(), first_ptr, first_copy.size() * sizeof(float), cudaMemcpyDeviceToHost));
    detector->detectAndCompute(img2, d2, whp, &desc1, true);  // a different frame through the same variable
    const bool fresh = desc1 != first_ptr;
    std::vector<float> again(first_copy.size());
    CHECK(cudaMemcpy(again.data(), first_ptr, again.size() * sizeof(float), cudaMemcpyDeviceToHost));
    const bool first_intact = again == first_copy;
    if (fresh) { CHECK(cudaFree(first_ptr)); }
    const float t1 = timer.read();
    for (int i = 0; i < nrepeats; i++) {
        detector->detectAndCompute(img1, d1, whp, &desc1, true);
        detector->detectAndCompute(img2, d2, whp, &desc2, true);
    }
    const float t2 = timer.read();
    for (int i = 0; i < nrepeats; i++) detector->match(d1, d2, desc1, desc2);
    const float t3 = timer.read();

    int good = 0;
    for (int i = 0; i < d1.num_pts; i++) good += d1.h_data[i].ambiguity < 0.8f;
    std::cout << "Number of features1: " << d1.num_pts << std::endl << "Number of features2: " << d2.num_pts << std::endl;
    std::cout << "Time for allocating image memory:  " << t0 << std::endl
              << "Time of detection and computation: " << (t2 - t1) / nrepeats << " (ms per pair)" << std::endl
              << "Time of matching surf keypoints:   " << (t3 - t2) / nrepeats << std::endl
              << "Matches with ambiguity < 0.8:      " << good << std::endl;
    if (d1.num_pts > 0)
        std::printf("first keypoint: x=%.3f y=%.3f scale=%.3f strength=%.3f laplace=%d match=%d score=%.4f\n", d1.h_data[0].x,
                    d1.h_data[0].y, d1.h_data[0].scale, d1.h_data[0].strength, d1.h_data[0].laplace, d1.h_data[0].match,
                    d1.h_data[0].score);

    std::printf("features1=%d\nfeatures2=%d\ngood=%d\ndesc_fresh=%d\nfirst_desc_intact=%d\n", d1.num_pts, d2.num_pts, good,
                fresh ? 1 : 0, first_intact ? 1 : 0);
    // per-row match results of the left image for the parity test: index and score
    if (const char* dump = std::getenv("SURF_DEMO_DUMP")) {
        std::ofstream o(dump, std::ios::binary);
        o.write((const char*)d1.h_data, sizeof(surf::SurfPoint) * (size_t)d1.num_pts);
    }
    surf::freeSurfData(d1);
    surf::freeSurfData(d2);
    CHECK(cudaFree(img1));
    CHECK(cudaFree(img2));
    if (desc1) cudaFree(desc1);
    if (desc2) cudaFree(desc2);
    return 0;
}
