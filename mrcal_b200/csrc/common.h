// Internal declarations shared by the host and device parts of libmrcal_b200.
// Nothing here is part of the C-ABI (that is include/mrcal_b200.h).
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/mrcal_b200.h"

namespace mb200 {

// State-vector scales: optimizer sees value/scale. Same constants as the
// reference's scales.h:40-48 (they define the meaning of b_packed, so they are
// part of the drop-in contract)
constexpr double kScaleFocal        = 500.0;
constexpr double kScaleCenter       = 20.0;
constexpr double kScaleRotCam       = 0.1 * M_PI / 180.0;
constexpr double kScaleTransCam     = 1.0;
constexpr double kScaleRotFrame     = 15.0 * M_PI / 180.0;
constexpr double kScaleTransFrame   = 1.0;
constexpr double kScalePoint        = kScaleTransFrame;
constexpr double kScaleWarp         = 0.01;
constexpr double kScaleDistortion   = 1.0;

void set_error(const char* fmt, ...);     // records + prints to stderr, like the reference's MSG()
const char* get_error();

inline bool sel_warp(mrcal_problem_selections_t s, int Nobs_board)
{ return s.do_optimize_calobject_warp && Nobs_board > 0; }

struct Dims
{
    int Ncam_i = 0, Ncam_e = 0, Nframes = 0, Npoints = 0, Npoints_fixed = 0;
    int Nobs_board = 0, Nobs_point = 0;
    int W = 0, H = 0;   // calibration object corners
    int Nmeas_tri = 0;  // measurements of the triangulated points: one per pair of observations of a point
};

// Everything integer about one problem: where each block of the state vector
// and of the measurement vector lives. One place computes it; the C-ABI layout
// functions and the kernels' launch parameters are views of this.
struct Layout
{
    Dims d;
    mrcal_lensmodel_t lensmodel;
    mrcal_problem_selections_t sel;   // normalised: core off if model has none, warp off if no boards

    int Nintr = 0;        // lens parameters per camera (incl. core)
    int Ncore_state = 0;  // 4 if the core is in the state
    int Ndist_state = 0;  // non-core parameters in the state
    int Nintr_state = 0;  // Ncore_state + Ndist_state
    bool splined = false;
    int  spline_order = 0, Nx = 0, Ny = 0;

    int i_intr0 = -1, i_extr0 = -1, i_frame0 = -1, i_point0 = -1, i_warp0 = -1;
    int Nstate = 0;
    int Npoints_variable = 0;

    // nonzeros per board measurement row: intrinsics part, and the full row for
    // a camera with / without extrinsics in the state
    int nnz_row_intr = 0;
    int nnz_row_board_geom = 0;   // frames + warp (no extrinsics)

    int m_board0 = 0, m_point0 = 0, m_tri0 = 0, m_reg0 = 0;
    int Nmeas_board = 0, Nmeas_point = 0, Nmeas_tri = 0, Nmeas_reg = 0, Nmeas = 0;
    int Nreg_dist = 0, Nreg_center = 0, Nreg_unity = 0;
};

bool make_layout(Layout* L, const Dims& d, mrcal_problem_selections_t sel, const mrcal_lensmodel_t* lensmodel);

// scale of each packed state element (b = value/scale)
void fill_state_scales(double* scale, const Layout& L);

}  // namespace mb200
