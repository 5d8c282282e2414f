// Host-side integer logic: lens-model descriptions and the layout of the state
// and measurement vectors. These define the MEANING of b_packed, x and J, so
// they must agree exactly with the reference (mrcal.c:29-882, 3290-3880); the
// tests compare every function here against the compiled reference over a grid
// of problem shapes and selections.
//
// Organisation differs from the reference: a table describes the lens models,
// and one Layout object (common.h) is computed per problem; the C entry points
// are thin views of it.
#include "common.h"

#include <cinttypes>
#include <cstdlib>

namespace mb200 {

static thread_local std::string g_error;

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    fprintf(stderr, "mrcal_b200: %s\n", buf);
}
const char* get_error() { return g_error.c_str(); }

struct ModelInfo
{
    mrcal_lensmodel_type_t type;
    const char* name;       // without configuration
    const char* name_tmpl;  // with "..." placeholders, as the reference reports it
    int  Nparams;           // <0: depends on the configuration
    bool has_config;
    bool behind_camera;
    bool noncentral;
};

// types.h:33-52 and the metadata at mrcal.c:252-288
static const ModelInfo kModels[] = {
    {MRCAL_LENSMODEL_PINHOLE,       "LENSMODEL_PINHOLE",       "LENSMODEL_PINHOLE",        4, false, false, false},
    {MRCAL_LENSMODEL_STEREOGRAPHIC, "LENSMODEL_STEREOGRAPHIC", "LENSMODEL_STEREOGRAPHIC",  4, false, true,  false},
    {MRCAL_LENSMODEL_LONLAT,        "LENSMODEL_LONLAT",        "LENSMODEL_LONLAT",         4, false, true,  false},
    {MRCAL_LENSMODEL_LATLON,        "LENSMODEL_LATLON",        "LENSMODEL_LATLON",         4, false, true,  false},
    {MRCAL_LENSMODEL_OPENCV4,       "LENSMODEL_OPENCV4",       "LENSMODEL_OPENCV4",        8, false, false, false},
    {MRCAL_LENSMODEL_OPENCV5,       "LENSMODEL_OPENCV5",       "LENSMODEL_OPENCV5",        9, false, false, false},
    {MRCAL_LENSMODEL_OPENCV8,       "LENSMODEL_OPENCV8",       "LENSMODEL_OPENCV8",       12, false, false, false},
    {MRCAL_LENSMODEL_OPENCV12,      "LENSMODEL_OPENCV12",      "LENSMODEL_OPENCV12",      16, false, false, false},
    {MRCAL_LENSMODEL_CAHVOR,        "LENSMODEL_CAHVOR",        "LENSMODEL_CAHVOR",         9, false, false, false},
    {MRCAL_LENSMODEL_CAHVORE,       "LENSMODEL_CAHVORE",       "LENSMODEL_CAHVORE_linearity=...", 12, true, false, true},
    {MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC, "LENSMODEL_SPLINED_STEREOGRAPHIC",
     "LENSMODEL_SPLINED_STEREOGRAPHIC_order=..._Nx=..._Ny=..._fov_x_deg=...", -1, true, true, false},
};
static const int kNmodels = (int)(sizeof(kModels) / sizeof(kModels[0]));

static const ModelInfo* model_info(mrcal_lensmodel_type_t type)
{
    for(int i = 0; i < kNmodels; i++)
        if(kModels[i].type == type) return &kModels[i];
    return nullptr;
}

bool make_layout(Layout* L, const Dims& d, mrcal_problem_selections_t sel, const mrcal_lensmodel_t* lensmodel)
{
    *L = Layout();
    L->d = d;
    L->lensmodel = *lensmodel;
    const ModelInfo* mi = model_info(lensmodel->type);
    if(mi == nullptr) { set_error("unknown lens model type %d", (int)lensmodel->type); return false; }

    L->Nintr = mrcal_lensmodel_num_params(lensmodel);
    if(L->Nintr < 4) { set_error("lens model has a bad parameter count %d", L->Nintr); return false; }
    if(d.Nobs_board <= 0) sel.do_optimize_calobject_warp = false;
    L->sel = sel;

    L->splined = lensmodel->type == MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC;
    if(L->splined)
    {
        const auto& c = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
        L->spline_order = c.order; L->Nx = c.Nx; L->Ny = c.Ny;
    }
    L->Ncore_state = sel.do_optimize_intrinsics_core        ? 4            : 0;
    L->Ndist_state = sel.do_optimize_intrinsics_distortions ? L->Nintr - 4 : 0;
    L->Nintr_state = L->Ncore_state + L->Ndist_state;
    L->Npoints_variable = d.Npoints - d.Npoints_fixed;

    int n = 0;
    const int Nsi = d.Ncam_i * L->Nintr_state;
    if(Nsi > 0) L->i_intr0 = 0;
    n += Nsi;
    if(sel.do_optimize_extrinsics && d.Ncam_e > 0) { L->i_extr0 = n; n += 6 * d.Ncam_e; }
    if(sel.do_optimize_frames)
    {
        if(d.Nframes > 0)          { L->i_frame0 = n; n += 6 * d.Nframes; }
        if(L->Npoints_variable > 0){ L->i_point0 = n; n += 3 * L->Npoints_variable; }
    }
    if(sel_warp(sel, d.Nobs_board)) { L->i_warp0 = n; n += 2; }
    L->Nstate = n;

    // mrcal.c:768-798
    if(L->splined)
    {
        const int run = L->spline_order + 1;
        L->nnz_row_intr = (sel.do_optimize_intrinsics_core ? 2 : 0) +
                          (sel.do_optimize_intrinsics_distortions ? run * run : 0);
    }
    else
        L->nnz_row_intr = L->Nintr_state - (sel.do_optimize_intrinsics_core ? 2 : 0);
    L->nnz_row_board_geom = (sel.do_optimize_frames ? 6 : 0) + (sel_warp(sel, d.Nobs_board) ? 2 : 0);

    L->Nmeas_board = d.Nobs_board > 0 ? d.Nobs_board * d.W * d.H * 2 : 0;
    L->Nmeas_point = d.Nobs_point * 2;
    if(sel.do_apply_regularization)
    {
        L->Nreg_dist   = d.Ncam_i * L->Ndist_state;
        L->Nreg_center = sel.do_optimize_intrinsics_core ? d.Ncam_i * 2 : 0;
    }
    L->Nreg_unity = (sel.do_apply_regularization_unity_cam01 && sel.do_optimize_extrinsics && d.Ncam_e > 0) ? 1 : 0;
    L->Nmeas_reg = L->Nreg_dist + L->Nreg_center + L->Nreg_unity;
    L->m_board0 = 0;
    L->m_point0 = L->Nmeas_board;
    L->m_tri0   = L->m_point0 + L->Nmeas_point;
    L->Nmeas_tri = d.Nmeas_tri;
    L->m_reg0   = L->m_tri0 + L->Nmeas_tri;
    L->Nmeas    = L->m_reg0 + L->Nmeas_reg;
    return true;
}

void fill_state_scales(double* scale, const Layout& L)
{
    int i = 0;
    for(int c = 0; c < L.d.Ncam_i; c++)
    {
        if(L.Ncore_state)
        {
            scale[i++] = kScaleFocal;  scale[i++] = kScaleFocal;
            scale[i++] = kScaleCenter; scale[i++] = kScaleCenter;
        }
        for(int k = 0; k < L.Ndist_state; k++) scale[i++] = kScaleDistortion;
    }
    if(L.i_extr0 >= 0)
        for(int c = 0; c < L.d.Ncam_e; c++)
        {
            for(int k = 0; k < 3; k++) scale[i++] = kScaleRotCam;
            for(int k = 0; k < 3; k++) scale[i++] = kScaleTransCam;
        }
    if(L.i_frame0 >= 0)
        for(int f = 0; f < L.d.Nframes; f++)
        {
            for(int k = 0; k < 3; k++) scale[i++] = kScaleRotFrame;
            for(int k = 0; k < 3; k++) scale[i++] = kScaleTransFrame;
        }
    if(L.i_point0 >= 0)
        for(int k = 0; k < 3 * L.Npoints_variable; k++) scale[i++] = kScalePoint;
    if(L.i_warp0 >= 0) { scale[i++] = kScaleWarp; scale[i++] = kScaleWarp; }
}

// A Layout for the pure-integer C entry points, which do not know W,H or the
// observation lists
static Layout quick_layout(int Ncam_i, int Ncam_e, int Nframes, int Npoints, int Npoints_fixed,
                           int Nobs_board, mrcal_problem_selections_t sel, const mrcal_lensmodel_t* lm)
{
    Dims d;
    d.Ncam_i = Ncam_i; d.Ncam_e = Ncam_e; d.Nframes = Nframes;
    d.Npoints = Npoints; d.Npoints_fixed = Npoints_fixed; d.Nobs_board = Nobs_board;
    Layout L;
    if(!make_layout(&L, d, sel, lm)) L.Nstate = -1;
    return L;
}

}  // namespace mb200

using namespace mb200;

////////////////////////////////////////////////////////////////////////////////
// Lens models
////////////////////////////////////////////////////////////////////////////////
extern "C" const char* mrcal_lensmodel_name_unconfigured(const mrcal_lensmodel_t* lensmodel)
{
    const ModelInfo* mi = model_info(lensmodel->type);
    return mi ? mi->name_tmpl : nullptr;
}

extern "C" bool mrcal_lensmodel_name(char* out, int size, const mrcal_lensmodel_t* lensmodel)
{
    const ModelInfo* mi = model_info(lensmodel->type);
    if(mi == nullptr) return false;
    int n;
    if(lensmodel->type == MRCAL_LENSMODEL_CAHVORE)
        n = snprintf(out, size, "%s_linearity=%.2f", mi->name, lensmodel->LENSMODEL_CAHVORE__config.linearity);
    else if(lensmodel->type == MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC)
    {
        const auto& c = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
        n = snprintf(out, size, "%s_order=%" PRIu16 "_Nx=%" PRIu16 "_Ny=%" PRIu16 "_fov_x_deg=%" PRIu16,
                     mi->name, c.order, c.Nx, c.Ny, c.fov_x_deg);
    }
    else
        n = snprintf(out, size, "%s", mi->name);
    return size > n;
}

extern "C" mrcal_lensmodel_type_t mrcal_lensmodel_type_from_name(const char* name)
{
    if(name == nullptr) return MRCAL_LENSMODEL_INVALID_TYPE;
    for(int i = 0; i < kNmodels; i++)
    {
        const size_t len = strlen(kModels[i].name);
        if(!kModels[i].has_config)
        {
            if(strcmp(name, kModels[i].name) == 0) return kModels[i].type;
        }
        else if(strncmp(name, kModels[i].name, len) == 0 && (name[len] == '\0' || name[len] == '_'))
            return kModels[i].type;
    }
    return MRCAL_LENSMODEL_INVALID_TYPE;
}

extern "C" bool mrcal_lensmodel_from_name(mrcal_lensmodel_t* lensmodel, const char* name)
{
    memset(lensmodel, 0, sizeof(*lensmodel));
    lensmodel->type = MRCAL_LENSMODEL_INVALID_TYPE;
    if(name == nullptr) return false;

    for(int i = 0; i < kNmodels; i++)
    {
        const ModelInfo& m = kModels[i];
        if(!m.has_config)
        {
            if(strcmp(name, m.name) == 0) { lensmodel->type = m.type; return true; }
            continue;
        }
        const size_t len = strlen(m.name);
        if(strncmp(name, m.name, len) != 0) continue;
        if(name[len] == '\0') { lensmodel->type = MRCAL_LENSMODEL_INVALID_MISSINGCONFIG; return false; }
        if(name[len] != '_') continue;

        // "NAME_key=value_key=value...", all keys required, in order, nothing trailing
        const char* cfg = name + len;
        int pos = -1;
        bool ok = false;
        if(m.type == MRCAL_LENSMODEL_CAHVORE)
        {
            double linearity;
            ok = 1 == sscanf(cfg, "_linearity=%lf%n", &linearity, &pos) && pos >= 0 && cfg[pos] == '\0';
            if(ok) lensmodel->LENSMODEL_CAHVORE__config.linearity = linearity;
        }
        else
        {
            uint16_t order, Nx, Ny, fov;
            ok = 4 == sscanf(cfg, "_order=%" SCNu16 "_Nx=%" SCNu16 "_Ny=%" SCNu16 "_fov_x_deg=%" SCNu16 "%n",
                             &order, &Nx, &Ny, &fov, &pos) && pos >= 0 && cfg[pos] == '\0';
            if(ok)
            {
                auto& c = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
                c.order = order; c.Nx = Nx; c.Ny = Ny; c.fov_x_deg = fov;
            }
        }
        if(ok) { lensmodel->type = m.type; return true; }
        memset(lensmodel, 0, sizeof(*lensmodel));
        lensmodel->type = MRCAL_LENSMODEL_INVALID_BADCONFIG;
        return false;
    }
    return false;
}

extern "C" mrcal_lensmodel_metadata_t mrcal_lensmodel_metadata(const mrcal_lensmodel_t* lensmodel)
{
    mrcal_lensmodel_metadata_t meta = {};
    const ModelInfo* mi = model_info(lensmodel->type);
    if(mi == nullptr) { set_error("unknown lens model %d", (int)lensmodel->type); return meta; }
    meta.has_core = true;
    meta.can_project_behind_camera = mi->behind_camera;
    meta.has_gradients = true;
    meta.noncentral = mi->noncentral;
    return meta;
}

extern "C" int mrcal_lensmodel_num_params(const mrcal_lensmodel_t* lensmodel)
{
    const ModelInfo* mi = model_info(lensmodel->type);
    if(mi == nullptr) return -1;
    if(mi->Nparams >= 0) return mi->Nparams;
    const auto& c = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
    return (int)c.Nx * (int)c.Ny * 2 + 4;   // two surfaces + core
}

extern "C" const char* const* mrcal_supported_lensmodel_names(void)
{
    static const char* names[kNmodels + 1];
    for(int i = 0; i < kNmodels; i++) names[i] = kModels[i].name_tmpl;
    names[kNmodels] = nullptr;
    return names;
}

namespace mb200 {
// mrcal.c:1904-1952: knots per unit of stereographic u
bool spline_segments_per_u(double* out, const mrcal_lensmodel_t* lm)
{
    const auto& c = lm->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
    int margin;
    if     (c.order == 2) { margin = 1; if(c.Nx < 3 || c.Ny < 3) { set_error("quadratic splines need Nx,Ny >= 3; got %d,%d", c.Nx, c.Ny); return false; } }
    else if(c.order == 3) { margin = 2; if(c.Nx < 4 || c.Ny < 4) { set_error("cubic splines need Nx,Ny >= 4; got %d,%d", c.Nx, c.Ny); return false; } }
    else { set_error("only spline order 2 and 3 are supported; got %d", c.order); return false; }
    const double th_edge_x = (double)c.fov_x_deg / 2. * M_PI / 180.;
    const double u_edge_x  = tan(th_edge_x / 2.) * 2;
    *out = (c.Nx - 1 - margin) / (u_edge_x * 2.);
    return true;
}
}  // namespace mb200

// replaces internal.h:85 / mrcal.c:1904-1965: the one derived quantity a lens model caches
extern "C" void _mrcal_precompute_lensmodel_data(mrcal_projection_precomputed_t* precomputed, const mrcal_lensmodel_t* lensmodel)
{
    if(lensmodel->type == MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC)
    {
        double v = 0.;
        if(!mb200::spline_segments_per_u(&v, lensmodel)) v = 0.;
        precomputed->LENSMODEL_SPLINED_STEREOGRAPHIC__precomputed.segments_per_u = v;
    }
    precomputed->ready = true;
}

extern "C" bool mrcal_knots_for_splined_models(double* ux, double* uy, const mrcal_lensmodel_t* lensmodel)
{
    if(lensmodel->type != MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC)
    {
        set_error("mrcal_knots_for_splined_models() works only with LENSMODEL_SPLINED_STEREOGRAPHIC");
        return false;
    }
    double spu;
    if(!spline_segments_per_u(&spu, lensmodel)) return false;
    const auto& c = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config;
    for(int i = 0; i < c.Nx; i++) ux[i] = ((double)i - (double)(c.Nx - 1) / 2.) / spu;
    for(int i = 0; i < c.Ny; i++) uy[i] = ((double)i - (double)(c.Ny - 1) / 2.) / spu;
    return true;
}

////////////////////////////////////////////////////////////////////////////////
// State layout
////////////////////////////////////////////////////////////////////////////////
#define QL() quick_layout(Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed, \
                          Nobservations_board, problem_selections, lensmodel)

extern "C" int mrcal_num_intrinsics_optimization_params(mrcal_problem_selections_t problem_selections,
                                                        const mrcal_lensmodel_t* lensmodel)
{
    return quick_layout(1, 0, 0, 0, 0, 0, problem_selections, lensmodel).Nintr_state;
}
extern "C" int mrcal_num_states_intrinsics(int Ncameras_intrinsics, mrcal_problem_selections_t problem_selections,
                                           const mrcal_lensmodel_t* lensmodel)
{
    return Ncameras_intrinsics * mrcal_num_intrinsics_optimization_params(problem_selections, lensmodel);
}
extern "C" int mrcal_num_states_extrinsics(int Ncameras_extrinsics, mrcal_problem_selections_t problem_selections)
{
    return problem_selections.do_optimize_extrinsics ? 6 * Ncameras_extrinsics : 0;
}
extern "C" int mrcal_num_states_frames(int Nframes, mrcal_problem_selections_t problem_selections)
{
    return problem_selections.do_optimize_frames ? 6 * Nframes : 0;
}
extern "C" int mrcal_num_states_points(int Npoints, int Npoints_fixed, mrcal_problem_selections_t problem_selections)
{
    return problem_selections.do_optimize_frames ? 3 * (Npoints - Npoints_fixed) : 0;
}
extern "C" int mrcal_num_states_calobject_warp(mrcal_problem_selections_t problem_selections, int Nobservations_board)
{
    return sel_warp(problem_selections, Nobservations_board) ? MRCAL_NSTATE_CALOBJECT_WARP : 0;
}
extern "C" int mrcal_num_states(int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                int Npoints, int Npoints_fixed, int Nobservations_board,
                                mrcal_problem_selections_t problem_selections, const mrcal_lensmodel_t* lensmodel)
{
    return mrcal_num_states_intrinsics(Ncameras_intrinsics, problem_selections, lensmodel) +
           mrcal_num_states_extrinsics(Ncameras_extrinsics, problem_selections) +
           mrcal_num_states_frames(Nframes, problem_selections) +
           mrcal_num_states_points(Npoints, Npoints_fixed, problem_selections) +
           mrcal_num_states_calobject_warp(problem_selections, Nobservations_board);
}

extern "C" int mrcal_state_index_intrinsics(int icam_intrinsics,
                                            int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                            int Npoints, int Npoints_fixed, int Nobservations_board,
                                            mrcal_problem_selections_t problem_selections,
                                            const mrcal_lensmodel_t* lensmodel)
{
    if(Ncameras_intrinsics <= 0) return -1;
    const int N = mrcal_num_intrinsics_optimization_params(problem_selections, lensmodel);
    if(N <= 0) return -1;
    if(icam_intrinsics < 0 || icam_intrinsics >= Ncameras_intrinsics) return -1;
    return icam_intrinsics * N;
}
extern "C" int mrcal_state_index_extrinsics(int icam_extrinsics,
                                            int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                            int Npoints, int Npoints_fixed, int Nobservations_board,
                                            mrcal_problem_selections_t problem_selections,
                                            const mrcal_lensmodel_t* lensmodel)
{
    if(Ncameras_extrinsics <= 0 || !problem_selections.do_optimize_extrinsics) return -1;
    if(icam_extrinsics < 0 || icam_extrinsics >= Ncameras_extrinsics) return -1;
    return mrcal_num_states_intrinsics(Ncameras_intrinsics, problem_selections, lensmodel) + 6 * icam_extrinsics;
}
extern "C" int mrcal_state_index_frames(int iframe,
                                        int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                        int Npoints, int Npoints_fixed, int Nobservations_board,
                                        mrcal_problem_selections_t problem_selections,
                                        const mrcal_lensmodel_t* lensmodel)
{
    if(Nframes <= 0 || !problem_selections.do_optimize_frames) return -1;
    if(iframe < 0 || iframe >= Nframes) return -1;
    return mrcal_num_states_intrinsics(Ncameras_intrinsics, problem_selections, lensmodel) +
           mrcal_num_states_extrinsics(Ncameras_extrinsics, problem_selections) + 6 * iframe;
}
extern "C" int mrcal_state_index_points(int i_point,
                                        int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                        int Npoints, int Npoints_fixed, int Nobservations_board,
                                        mrcal_problem_selections_t problem_selections,
                                        const mrcal_lensmodel_t* lensmodel)
{
    const int Nvar = Npoints - Npoints_fixed;
    if(Nvar <= 0 || !problem_selections.do_optimize_frames) return -1;
    if(i_point < 0 || i_point >= Nvar) return -1;
    return mrcal_num_states_intrinsics(Ncameras_intrinsics, problem_selections, lensmodel) +
           mrcal_num_states_extrinsics(Ncameras_extrinsics, problem_selections) +
           mrcal_num_states_frames(Nframes, problem_selections) + 3 * i_point;
}
extern "C" int mrcal_state_index_calobject_warp(int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                                int Npoints, int Npoints_fixed, int Nobservations_board,
                                                mrcal_problem_selections_t problem_selections,
                                                const mrcal_lensmodel_t* lensmodel)
{
    if(!sel_warp(problem_selections, Nobservations_board)) return -1;
    return mrcal_num_states_intrinsics(Ncameras_intrinsics, problem_selections, lensmodel) +
           mrcal_num_states_extrinsics(Ncameras_extrinsics, problem_selections) +
           mrcal_num_states_frames(Nframes, problem_selections) +
           mrcal_num_states_points(Npoints, Npoints_fixed, problem_selections);
}

////////////////////////////////////////////////////////////////////////////////
// Measurement layout
////////////////////////////////////////////////////////////////////////////////
extern "C" int mrcal_num_measurements_boards(int Nobservations_board,
                                             int calibration_object_width_n, int calibration_object_height_n)
{
    if(Nobservations_board <= 0) return 0;
    return Nobservations_board * calibration_object_width_n * calibration_object_height_n * 2;
}
extern "C" int mrcal_measurement_index_boards(int i_observation_board,
                                              int Nobservations_board, int Nobservations_point,
                                              int calibration_object_width_n, int calibration_object_height_n)
{
    if(Nobservations_board <= 0) return -1;
    return mrcal_num_measurements_boards(i_observation_board, calibration_object_width_n, calibration_object_height_n);
}
extern "C" int mrcal_num_measurements_points(int Nobservations_point) { return Nobservations_point * 2; }
extern "C" int mrcal_measurement_index_points(int i_observation_point,
                                              int Nobservations_board, int Nobservations_point,
                                              int calibration_object_width_n, int calibration_object_height_n)
{
    if(Nobservations_point <= 0) return -1;
    return mrcal_num_measurements_boards(Nobservations_board, calibration_object_width_n, calibration_object_height_n) +
           i_observation_point * 2;
}

// Observations of one triangulated point are stored consecutively, the last
// flagged last_in_set; a set of n observations yields n(n-1)/2 measurements
// (mrcal.c:485-524)
extern "C" int mrcal_num_measurements_points_triangulated_initial_Npoints(
        const mrcal_observation_point_triangulated_t* obs, int Nobs, int Npoints)
{
    if(obs == nullptr || Nobs <= 0) return 0;
    int Nmeas = 0, ipoint = 0, i = 0;
    while(i < Nobs && (Npoints < 0 || ipoint < Npoints))
    {
        int n = 1;
        while(i < Nobs && !obs[i].last_in_set) { n++; i++; }
        Nmeas += n * (n - 1) / 2;
        ipoint++;
        i++;
    }
    return Nmeas;
}
extern "C" int mrcal_num_measurements_points_triangulated(const mrcal_observation_point_triangulated_t* obs, int Nobs)
{
    return mrcal_num_measurements_points_triangulated_initial_Npoints(obs, Nobs, -1);
}
extern "C" int mrcal_measurement_index_points_triangulated(int i_point_triangulated,
                                                           int Nobservations_board, int Nobservations_point,
                                                           const mrcal_observation_point_triangulated_t* obs, int Nobs,
                                                           int calibration_object_width_n, int calibration_object_height_n)
{
    if(obs == nullptr || Nobs <= 0) return -1;
    return mrcal_num_measurements_boards(Nobservations_board, calibration_object_width_n, calibration_object_height_n) +
           mrcal_num_measurements_points(Nobservations_point) +
           mrcal_num_measurements_points_triangulated_initial_Npoints(obs, Nobs, i_point_triangulated);
}
extern "C" int mrcal_num_measurements_regularization(int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                                     int Npoints, int Npoints_fixed, int Nobservations_board,
                                                     mrcal_problem_selections_t problem_selections,
                                                     const mrcal_lensmodel_t* lensmodel)
{
    // Follows the reference exactly, including that the centre-pixel terms are
    // counted off do_optimize_intrinsics_core as given (mrcal.c:382-393, 676-690)
    int per_camera = 0;
    if(problem_selections.do_apply_regularization)
    {
        if(problem_selections.do_optimize_intrinsics_distortions)
            per_camera += mrcal_lensmodel_num_params(lensmodel) - 4;
        if(problem_selections.do_optimize_intrinsics_core)
            per_camera += 2;
    }
    return Ncameras_intrinsics * per_camera +
           ((problem_selections.do_apply_regularization_unity_cam01 &&
             problem_selections.do_optimize_extrinsics && Ncameras_extrinsics > 0) ? 1 : 0);
}
extern "C" int mrcal_measurement_index_regularization(
        const mrcal_observation_point_triangulated_t* obs_tri, int Nobs_tri,
        int calibration_object_width_n, int calibration_object_height_n,
        int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
        int Npoints, int Npoints_fixed, int Nobservations_board, int Nobservations_point,
        mrcal_problem_selections_t problem_selections, const mrcal_lensmodel_t* lensmodel)
{
    if(mrcal_num_measurements_regularization(Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                             Nobservations_board, problem_selections, lensmodel) <= 0)
        return -1;
    return mrcal_num_measurements_boards(Nobservations_board, calibration_object_width_n, calibration_object_height_n) +
           mrcal_num_measurements_points(Nobservations_point) +
           mrcal_num_measurements_points_triangulated(obs_tri, Nobs_tri);
}
extern "C" int mrcal_num_measurements(int Nobservations_board, int Nobservations_point,
                                      const mrcal_observation_point_triangulated_t* obs_tri, int Nobs_tri,
                                      int calibration_object_width_n, int calibration_object_height_n,
                                      int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                      int Npoints, int Npoints_fixed,
                                      mrcal_problem_selections_t problem_selections, const mrcal_lensmodel_t* lensmodel)
{
    return mrcal_num_measurements_boards(Nobservations_board, calibration_object_width_n, calibration_object_height_n) +
           mrcal_num_measurements_points(Nobservations_point) +
           mrcal_num_measurements_points_triangulated(obs_tri, Nobs_tri) +
           mrcal_num_measurements_regularization(Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                                                 Nobservations_board, problem_selections, lensmodel);
}

extern "C" int _mrcal_num_j_nonzero(int Nobservations_board, int Nobservations_point,
                                    const mrcal_observation_point_triangulated_t* obs_tri, int Nobs_tri,
                                    int calibration_object_width_n, int calibration_object_height_n,
                                    int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                    int Npoints, int Npoints_fixed,
                                    const mrcal_observation_board_t* observations_board,
                                    const mrcal_observation_point_t* observations_point,
                                    mrcal_problem_selections_t problem_selections, const mrcal_lensmodel_t* lensmodel)
{
    // The reference counts from the selections AS GIVEN (mrcal.c:743-882); it
    // only gates the warp on there being boards. Do the same
    const mrcal_problem_selections_t s = problem_selections;
    const bool splined = lensmodel->type == MRCAL_LENSMODEL_SPLINED_STEREOGRAPHIC;
    int Nintr_row;
    if(splined)
    {
        const int run = lensmodel->LENSMODEL_SPLINED_STEREOGRAPHIC__config.order + 1;
        Nintr_row = (s.do_optimize_intrinsics_core ? 4 : 0) + (s.do_optimize_intrinsics_distortions ? run * run : 0);
    }
    else
        Nintr_row = mrcal_num_intrinsics_optimization_params(s, lensmodel);
    if(s.do_optimize_intrinsics_core) Nintr_row -= 2;   // x sees fx,cx only; y sees fy,cy only

    const bool warp = sel_warp(s, Nobservations_board);
    long N = (long)Nobservations_board *
             ((s.do_optimize_frames ? 6 : 0) + (s.do_optimize_extrinsics ? 6 : 0) + (warp ? 2 : 0) + Nintr_row);
    if(s.do_optimize_extrinsics)
        for(int i = 0; i < Nobservations_board; i++)
            if(observations_board[i].icam.extrinsics < 0) N -= 6;
    N *= 2 * calibration_object_width_n * calibration_object_height_n;

    for(int i = 0; i < Nobservations_point; i++)
    {
        N += 2 * Nintr_row;
        if(s.do_optimize_frames && observations_point[i].i_point < Npoints - Npoints_fixed) N += 2 * 3;
        if(s.do_optimize_extrinsics && observations_point[i].icam.extrinsics >= 0)          N += 2 * 6;
    }

    if(obs_tri != nullptr && Nobs_tri > 0)
        for(int i0 = 0; i0 < Nobs_tri; i0++)
        {
            if(obs_tri[i0].last_in_set) continue;
            const int Nvars0 = Nintr_row + ((s.do_optimize_extrinsics && obs_tri[i0].icam.extrinsics >= 0) ? 6 : 0);
            int i1 = i0;
            do
            {
                i1++;
                const int Nvars1 = Nintr_row + ((s.do_optimize_extrinsics && obs_tri[i1].icam.extrinsics >= 0) ? 6 : 0);
                N += Nvars0 + Nvars1;
            } while(!obs_tri[i1].last_in_set);
        }

    int reg_per_camera = 0;
    if(s.do_apply_regularization)
    {
        if(s.do_optimize_intrinsics_distortions) reg_per_camera += mrcal_lensmodel_num_params(lensmodel) - 4;
        if(s.do_optimize_intrinsics_core)        reg_per_camera += 2;
    }
    if(splined)
    {
        if(s.do_apply_regularization)
        {
            // each knot row touches both surfaces; centre-pixel rows touch one value
            N += (long)Ncameras_intrinsics * 2 * reg_per_camera;
            if(s.do_optimize_intrinsics_core) N -= Ncameras_intrinsics * 2;
        }
    }
    else
        N += (long)Ncameras_intrinsics * reg_per_camera;

    if(s.do_apply_regularization_unity_cam01 && s.do_optimize_extrinsics && Ncameras_extrinsics > 0)
        N += 3;
    return (int)N;
}

////////////////////////////////////////////////////////////////////////////////
// pack / unpack
////////////////////////////////////////////////////////////////////////////////
static void scale_state_vector(double* b, bool pack,
                               int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                               int Npoints, int Npoints_fixed, int Nobservations_board,
                               mrcal_problem_selections_t problem_selections, const mrcal_lensmodel_t* lensmodel)
{
    Layout L = QL();
    if(L.Nstate <= 0) return;
    std::vector<double> scale(L.Nstate);
    fill_state_scales(scale.data(), L);
    if(pack) for(int i = 0; i < L.Nstate; i++) b[i] /= scale[i];
    else     for(int i = 0; i < L.Nstate; i++) b[i] *= scale[i];
}
extern "C" void mrcal_pack_solver_state_vector(double* b,
                                               int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                               int Npoints, int Npoints_fixed, int Nobservations_board,
                                               mrcal_problem_selections_t problem_selections,
                                               const mrcal_lensmodel_t* lensmodel)
{
    scale_state_vector(b, true, Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                       Nobservations_board, problem_selections, lensmodel);
}
extern "C" void mrcal_unpack_solver_state_vector(double* b,
                                                 int Ncameras_intrinsics, int Ncameras_extrinsics, int Nframes,
                                                 int Npoints, int Npoints_fixed, int Nobservations_board,
                                                 mrcal_problem_selections_t problem_selections,
                                                 const mrcal_lensmodel_t* lensmodel)
{
    scale_state_vector(b, false, Ncameras_intrinsics, Ncameras_extrinsics, Nframes, Npoints, Npoints_fixed,
                       Nobservations_board, problem_selections, lensmodel);
}

// In a vanilla calibration (stationary cameras) each intrinsics index pairs with
// exactly one extrinsics index (-1: the reference camera). mrcal.c:3893-3976
extern "C" bool mrcal_corresponding_icam_extrinsics(int* icam_extrinsics, int icam_intrinsics,
                                                    int Ncameras_intrinsics, int Ncameras_extrinsics,
                                                    int Nobservations_board,
                                                    const mrcal_observation_board_t* observations_board,
                                                    int Nobservations_point,
                                                    const mrcal_observation_point_t* observations_point)
{
    if(!(Ncameras_intrinsics == Ncameras_extrinsics || Ncameras_intrinsics == Ncameras_extrinsics + 1))
    {
        set_error("cannot compute icam_extrinsics: not a vanilla calibration problem (stationary cameras, cam0 is reference)");
        return false;
    }
    const int kUnset = -100;
    std::vector<int> to_e(Ncameras_intrinsics, kUnset), to_i(Ncameras_extrinsics + 1, kUnset);
    auto see = [&](const mrcal_camera_index_t& icam, int i, const char* what) -> bool
    {
        const int ci = icam.intrinsics;
        const int ce = icam.extrinsics < 0 ? -1 : icam.extrinsics;
        if(ci < 0 || ci >= Ncameras_intrinsics || ce >= Ncameras_extrinsics)
        {
            set_error("%s observation %d has out-of-range camera indices %d,%d", what, i, ci, ce);
            return false;
        }
        if(to_i[ce + 1] == kUnset) to_i[ce + 1] = ci;
        else if(to_i[ce + 1] != ci)
        {
            set_error("cannot compute icam_extrinsics: %s observation %d pairs (%d,%d) but (%d,%d) was seen before",
                      what, i, ci, ce, to_i[ce + 1], ce);
            return false;
        }
        if(to_e[ci] == kUnset) to_e[ci] = ce;
        else if(to_e[ci] != ce)
        {
            set_error("cannot compute icam_extrinsics: %s observation %d pairs (%d,%d) but (%d,%d) was seen before",
                      what, i, ci, ce, ci, to_e[ci]);
            return false;
        }
        return true;
    };
    for(int i = 0; i < Nobservations_board; i++) if(!see(observations_board[i].icam, i, "board")) return false;
    for(int i = 0; i < Nobservations_point; i++) if(!see(observations_point[i].icam, i, "point")) return false;
    if(icam_intrinsics < 0 || icam_intrinsics >= Ncameras_intrinsics) return false;
    *icam_extrinsics = to_e[icam_intrinsics];
    return true;
}
