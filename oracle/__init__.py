"""CPU oracle for the pressure-surrogate hot path  --  TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy/SciPy, the algorithm of the reference
(pauloacs/Solving-Poisson-s-Equation-through-DL-for-CFD-apllications) for the
per-timestep surrogate prediction.  Every function cites the reference
file:line it follows.  Path aliases (all relative to the reference root):

  SMC  = Improved_SM/deltaU_to_deltaP/source/pressureSM_deltas/SM_call.py
  UTL  = Improved_SM/deltaU_to_deltaP/source/pressureSM_deltas/utils.py
  NNS  = Improved_SM/deltaU_to_deltaP/source/pressureSM_deltas/NNs.py
  GRAD = Improved_SM/U_to_gradP/evaluation/Eval_dual_Dense_onlycil.py
  PMP  = Thesis_Work/Chapter5/parallelized/test_case/python_module.py

It is the CHECKER, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under
``solving-poisson-s-equation-through-dl-for-cfd-apllications_b200/`` imports it,
and the product path raises if the CUDA library is missing.

Parity pin: the reference holds no golden vectors.  The
oracle is pinned instead against OUTPUTS OF THE REFERENCE ITSELF, produced in
the build container by ``tests/golden/make_golden.py``: that script imports the
reference modules from /root/reference (with stand-ins for the third-party
packages that are not installable here: tensorflow, h5py, shapely, matplotlib,
mpi4py) and records what ``Evaluation.computeOnlyOnce`` / ``timeStep`` /
``assemble_prediction`` (SMC, GRAD) and ``init_func`` / ``py_func`` (PMP)
return on seeded synthetic inputs.  ``tests/test_oracle_golden.py`` checks this
package against those fixtures.

Substitutions relative to the reference:
  * Keras Dense stack  -> float32 NumPy ``x @ W + b`` / ReLU   (NNS:24-33)
  * shapely convex_hull -> scipy.spatial.ConvexHull             (SMC:128-133)
  * mpl Path.contains_points -> convex point-in-polygon test    (SMC:135-136)
  * pca.transform -> (X - mean_) @ components_.T                (PMP:349)
  * HDF5 frames -> in-memory arrays
"""
