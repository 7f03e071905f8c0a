// s2mv_image — headless counterpart of the reference's image driver (image_io.cpp): same 16
// arguments in the same order, the same sequence of stage calls (image_io.cpp:171-292) through the
// reference's own function names (include/s2mv_compat.h: the symbols libs2mv.so exports for an
// unchanged image_io.cpp), results written to files instead of an OpenCV window.
//
//   s2mv_image <left> <right> <ad coeff> <census coeff> <ndisp> <zerodisp> <upper color limit>
//              <lower color limit> <upper spatial limit> <lower spatial limit> <num views> <angle>
//              <out width> <out height> <thresh s> <thresh h>
//
// <left>/<right> are names under ./img/ without extension, as in the reference, or paths to .bmp files.
// Output directory: $S2MV_OUT (default ./out): disp_l.bmp, disp_r.bmp, disp_l.f32, disp_r.f32,
// view_<k>.bmp, interlaced.bmp.
#include <stdlib.h>

#ifdef S2MV_REFERENCE_HEADERS
// Boundary proof (oracle/build_ref.sh -> oracle/_ref/s2mv_image_refhdr): this same file compiled against the
// REFERENCE's own headers -- the include list of image_io.cpp:10-26, OpenCV satisfied by empty stubs -- and
// linked against libs2mv.so.  If a prototype here differed from the reference's, this would not link.
#include "cuda_utils.h"
#include "d_filter_gaussian.h"
#include "d_filter.h"
#include "d_filter_bilateral.h"
#include "d_dibr_occl.h"
#include "d_dibr_fwarp.h"
#include "d_dibr_bwarp.h"
#include "d_dc_wta.h"
#include "d_dc_hslo.h"
#include "d_dr_dcc.h"
#include "d_dr_irv.h"
#include "d_ca_cross.h"
#include "d_ci_adcensus.h"
#include "d_ci_census.h"
#include "d_ci_ad.h"
#include "d_tx_scale.h"
#include "d_mux_multiview.h"
#else
#include "../include/s2mv_compat.h"
#endif
#include "bmp_io.h"

static std::string image_path(const char *arg)
{
    std::string s(arg);
    if (s.find('/') != std::string::npos || (s.size() > 4 && s.substr(s.size() - 4) == ".bmp")) return s;
    return "./img/" + s + ".bmp";  // image_io.cpp:79-90
}

int main(int argc, char **argv)
{
    if (argc != 17) {
        printf("Place images in img subdir: \nthen input file names directly w/o dir extension \n");
        printf("Usage: ./program [left file] [right file] [ad coeff] [census coeff] [ndisp] [zerodisp] [upper color limit] "
               "[lower color limit] [upper spatial limit] [lower spatial limit] [num views] [angle] [out width] [out height] "
               "[thresh s] [thresh h]\n");
        return -1;
    }
    std::vector<uint8_t> img_l, img_r;
    int num_rows = 0, num_cols = 0, r2 = 0, c2 = 0;
    if (!bmpio::read_bmp(image_path(argv[1]), img_l, num_rows, num_cols) ||
        !bmpio::read_bmp(image_path(argv[2]), img_r, r2, c2) || r2 != num_rows || c2 != num_cols) {
        printf("Error! Could not read image files from disk! \n");
        return -1;
    }
    const int elem_sz = 3;
    // image_io.cpp:117-130
    float ad_coeff = atof(argv[3]), census_coeff = atof(argv[4]);
    int num_disp = atoi(argv[5]), zero_disp = atoi(argv[6]);
    float ucd = atof(argv[7]), lcd = atof(argv[8]);
    int usd = atoi(argv[9]), lsd = atoi(argv[10]);
    int num_views = atoi(argv[11]);
    float angle = atof(argv[12]);
    int num_cols_out = atoi(argv[13]), num_rows_out = atoi(argv[14]);
    int thresh_s = atoi(argv[15]);
    float thresh_h = atof(argv[16]);
    if (num_disp < 1 || num_views < 2 || num_views > 16 || num_cols_out < 1 || num_rows_out < 1) {
        printf("Error! Parameters out of range\n");
        return -1;
    }
    printf("Input Width:             %d\nInput Height:            %d\nNumber of Views:         %d\n", num_cols, num_rows, num_views);
    printf("Number of Disparities:   %d\nZero Disparity Index:    %d\n", num_disp, zero_disp);

    const size_t n = (size_t)num_rows * num_cols;
    auto planes = [&](std::vector<float> &store, std::vector<float *> &tab, int count) {
        store.assign(n * count, 0.f);
        tab.resize(count);
        for (int d = 0; d < count; ++d) tab[d] = store.data() + (size_t)d * n;
    };
    std::vector<float> s_cl, s_cr, s_al, s_ar;
    std::vector<float *> cost_l, cost_r, acost_l, acost_r;
    planes(s_cl, cost_l, num_disp); planes(s_cr, cost_r, num_disp);
    planes(s_al, acost_l, num_disp); planes(s_ar, acost_r, num_disp);
    std::vector<uint8_t> s_xl(4 * n), s_xr(4 * n);
    unsigned char *cross_l[4], *cross_r[4];
    for (int a = 0; a < 4; ++a) { cross_l[a] = s_xl.data() + a * n; cross_r[a] = s_xr.data() + a * n; }

    // image_io.cpp:171-223
    ci_adcensus(img_l.data(), img_r.data(), cost_l.data(), cost_r.data(), ad_coeff, census_coeff, num_disp, zero_disp,
                num_rows, num_cols, elem_sz);
    ca_cross(img_l.data(), cross_l, cost_l.data(), acost_l.data(), ucd, lcd, usd, lsd, num_disp, num_rows, num_cols, elem_sz);
    ca_cross(img_r.data(), cross_r, cost_r.data(), acost_r.data(), ucd, lcd, usd, lsd, num_disp, num_rows, num_cols, elem_sz);
    std::vector<float> disp_l(n), disp_r(n);
    dc_wta(acost_l.data(), disp_l.data(), num_disp, zero_disp, num_rows, num_cols);
    dc_wta(acost_r.data(), disp_r.data(), num_disp, zero_disp, num_rows, num_cols);
    // image_io.cpp:235-243
    std::vector<uint8_t> outl_l(n, 0), outl_r(n, 0);
    dr_dcc(outl_l.data(), outl_r.data(), disp_l.data(), disp_r.data(), num_rows, num_cols);
    dr_irv(disp_l.data(), outl_l.data(), cross_l, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd, 1);
    dr_irv(disp_r.data(), outl_r.data(), cross_r, thresh_s, thresh_h, num_rows, num_cols, num_disp, zero_disp, usd, 1);
    filter_bilateral_1(disp_l.data(), 7, 7, 7, num_rows, num_cols, num_disp);
    filter_bilateral_1(disp_r.data(), 7, 7, 7, num_rows, num_cols, num_disp);
    // image_io.cpp:255-292
    std::vector<uint8_t> occl_l(n, 0), occl_r(n, 0);
    dibr_occl(occl_l.data(), occl_r.data(), disp_l.data(), disp_r.data(), num_rows, num_cols);
    filter_bleed_1(occl_l.data(), 1, num_rows, num_cols);
    filter_bleed_1(occl_r.data(), 1, num_rows, num_cols);
    std::vector<float> mask_l(n), mask_r(n);
    dibr_occl_to_mask(mask_l.data(), mask_r.data(), occl_l.data(), occl_r.data(), num_rows, num_cols);
    std::vector<std::vector<uint8_t>> views(num_views);
    std::vector<unsigned char *> vtab(num_views);
    views[0] = img_r;
    views[num_views - 1] = img_l;
    for (int v = 1; v < num_views - 1; ++v) views[v].assign(n * 3, 0);
    for (int v = 0; v < num_views; ++v) vtab[v] = views[v].data();
    for (int v = 1; v < num_views - 1; ++v) {
        float shift = 1.0 - ((1.0 * (float)v) / ((float)num_views - 1.0));  // image_io.cpp:281
        dibr_dbm(vtab[v], img_l.data(), img_r.data(), disp_l.data(), disp_r.data(), occl_l.data(), occl_r.data(),
                 mask_l.data(), mask_r.data(), shift, num_rows, num_cols, elem_sz);
    }
    std::vector<uint8_t> mux((size_t)num_rows_out * num_cols_out * 3);
    mux_multiview(vtab.data(), mux.data(), num_views, angle, num_rows, num_cols, num_rows_out, num_cols_out, elem_sz);

    const std::string out = getenv("S2MV_OUT") ? getenv("S2MV_OUT") : "./out";
    bool ok = bmpio::write_plane_bmp(out + "/disp_l.bmp", disp_l.data(), num_rows, num_cols) &&
              bmpio::write_plane_bmp(out + "/disp_r.bmp", disp_r.data(), num_rows, num_cols) &&
              bmpio::write_raw(out + "/disp_l.f32", disp_l.data(), n * sizeof(float)) &&
              bmpio::write_raw(out + "/disp_r.f32", disp_r.data(), n * sizeof(float)) &&
              bmpio::write_bmp(out + "/interlaced.bmp", mux.data(), num_rows_out, num_cols_out);
    for (int v = 0; v < num_views && ok; ++v)
        ok = bmpio::write_bmp(out + "/view_" + std::to_string(v) + ".bmp", vtab[v], num_rows, num_cols);
    if (!ok) { printf("Error! Could not write results to %s (does the directory exist?)\n", out.c_str()); return -1; }
    printf("Wrote disparities, %d views and the interlaced frame to %s\n", num_views, out.c_str());
    return 0;
}
