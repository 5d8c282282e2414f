import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["MRCAL_B200_DEBUG_OUTLIERS"] = "1"
import numpy as np, problems, mrcal_b200
kw = dict(problems.solve_cases())["tri_divergent_rejection"]
ref = np.load(os.path.join(os.path.dirname(__file__), "tmp_tri_div.npz"))
k0 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
k0["do_apply_outlier_rejection"] = False
r = mrcal_b200.optimize(**k0)
print("norej: norm2", float(r["x"] @ r["x"]), "db", np.abs(r["b_packed"] - ref["b0"]).max(), "dx", np.abs(r["x"] - ref["x0"]).max())
k1 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
r = mrcal_b200.optimize(**k1)
print("rej: norm2", float(r["x"] @ r["x"]), r["Noutliers_triangulated_point"], np.flatnonzero(k1["observations_point_triangulated"][:, 0] != kw["observations_point_triangulated"][:, 0]))
print(k1["indices_point_triangulated_camintrinsics_camextrinsics"][:12])
