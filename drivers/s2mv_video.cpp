// s2mv_video — headless counterpart of the reference's video driver (video_io.cpp): the same 15
// arguments in the same order, every frame through `adcensus_stm` (video_io.cpp:158; the symbol
// libs2mv.so exports for an unchanged video_io.cpp).  OpenCV's VideoCapture is not available in this
// image, so the stream is raw side-by-side BGR24 frames:
//
//   s2mv_video <file> <num views> <angle> <out width> <out height> <ndisp> <zerodisp> <ad coeff>
//              <census coeff> <upper color limit> <lower color limit> <upper spatial limit>
//              <lower spatial limit> <thresh s> <thresh h>
//
// <file>: a name under ./vid/ as in the reference, or a path.  Frame geometry comes from
// $S2MV_SBS_COLS x $S2MV_ROWS (a raw stream has no header); a .bmp file is read as a one-frame stream.
// Output ($S2MV_OUT, default ./out): interlaced.bgr (raw frames), disp_l.f32 / disp_r.f32 (raw planes,
// frame after frame) and the per-frame wall time the reference prints (video_io.cpp:160-162).
#include <stdlib.h>
#include <time.h>

#ifdef S2MV_REFERENCE_HEADERS
// Boundary proof (oracle/build_ref.sh -> oracle/_ref/s2mv_video_refhdr): this same file compiled against the
// REFERENCE's own d_io.h (video_io.cpp:11-14; OpenCV satisfied by empty stubs) and linked against libs2mv.so.
#include "cuda_utils.h"
#include "d_io.h"
#else
#include "../include/s2mv_compat.h"
#endif
#include "bmp_io.h"

static double now_s()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc != 16) {
        printf("Place videos in vid subdir: \nthen input file name directly \n");
        printf("Usage: ./program [file] [num views] [angle] [out width] [out height] [ndisp] [zerodisp] [ad coeff] "
               "[census coeff] [upper color limit] [lower color limit] [upper spatial limit] [lower spatial limit] "
               "[thresh s] [thresh h]\n");
        return -1;
    }
    std::string path(argv[1]);
    if (path.find('/') == std::string::npos) path = "./vid/" + path;  // video_io.cpp:67-70
    // video_io.cpp:96-110
    int num_views = atoi(argv[2]);
    float angle = atof(argv[3]);
    int num_cols_out = atoi(argv[4]), num_rows_out = atoi(argv[5]);
    int num_disp = atoi(argv[6]), zero_disp = atoi(argv[7]);
    float ad_coeff = atof(argv[8]), census_coeff = atof(argv[9]);
    float ucd = atof(argv[10]), lcd = atof(argv[11]);
    int usd = atoi(argv[12]), lsd = atoi(argv[13]);
    int thresh_s = atoi(argv[14]);
    float thresh_h = atof(argv[15]);

    int num_rows = 0, num_cols_sbs = 0;
    std::vector<uint8_t> frame;
    FILE *f = nullptr;
    const bool is_bmp = path.size() > 4 && path.substr(path.size() - 4) == ".bmp";
    if (is_bmp) {
        if (!bmpio::read_bmp(path, frame, num_rows, num_cols_sbs)) { printf("Video cannot be read!\nAborting...\n"); return -1; }
    } else {
        num_cols_sbs = getenv("S2MV_SBS_COLS") ? atoi(getenv("S2MV_SBS_COLS")) : 0;
        num_rows = getenv("S2MV_ROWS") ? atoi(getenv("S2MV_ROWS")) : 0;
        f = fopen(path.c_str(), "rb");
        if (!f || num_cols_sbs < 2 || num_rows < 1) { printf("Video cannot be read!\nAborting...\n"); return -1; }
        frame.resize((size_t)num_rows * num_cols_sbs * 3);
    }
    const int num_cols = num_cols_sbs / 2, elem_sz = 3;  // video_io.cpp:88-89
    printf("Input Width (SBS):       %d\nInput Width (Single):    %d\nInput Height:            %d\n", num_cols_sbs, num_cols, num_rows);
    if (num_disp < 1 || num_views < 2 || num_views > 16 || num_cols_out < 1 || num_rows_out < 1) {
        printf("Error! Parameters out of range\n");
        return -1;
    }
    const size_t n = (size_t)num_rows * num_cols;
    std::vector<float> disp_l(n), disp_r(n);
    std::vector<uint8_t> interlaced((size_t)num_rows_out * num_cols_out * 3);
    const std::string out = getenv("S2MV_OUT") ? getenv("S2MV_OUT") : "./out";
    FILE *fo = fopen((out + "/interlaced.bgr").c_str(), "wb"), *fl = fopen((out + "/disp_l.f32").c_str(), "wb"),
         *fr = fopen((out + "/disp_r.f32").c_str(), "wb");
    if (!fo || !fl || !fr) { printf("Error! Could not open outputs in %s (does the directory exist?)\n", out.c_str()); return -1; }
    int frames = 0;
    double total = 0;
    for (;;) {
        if (!is_bmp && fread(frame.data(), 1, frame.size(), f) != frame.size()) break;
        const double t0 = now_s();
        adcensus_stm(frame.data(), disp_l.data(), disp_r.data(), interlaced.data(), num_rows, num_cols_sbs, num_cols,
                     num_rows_out, num_cols_out, elem_sz, num_views, (int)angle, num_disp, zero_disp, ad_coeff, census_coeff,
                     ucd, lcd, usd, lsd, thresh_s, thresh_h);
        const double dt = now_s() - t0;
        total += dt;
        printf("Frame %d: %f ms\n", frames, dt * 1e3);
        fwrite(interlaced.data(), 1, interlaced.size(), fo);
        fwrite(disp_l.data(), sizeof(float), n, fl);
        fwrite(disp_r.data(), sizeof(float), n, fr);
        ++frames;
        if (is_bmp) break;
    }
    fclose(fo); fclose(fl); fclose(fr);
    if (f) fclose(f);
    printf("%d frame(s), mean %f ms/frame\n", frames, frames ? total / frames * 1e3 : 0.0);
    return frames > 0 ? 0 : -1;
}
