/* solid_capi.h -- C API of libfsilbm_solid.so: the CPU-only structural side of the stand-in driver (beam_solver.cpp,
 * solid_body.cpp) for the Python tests and for bench.py's flexible-plate workload.  This is NOT part of the drop-in
 * boundary (include/fsilbm.h): in the reference these procedures live in the Fortran driver (Solidbody.f90 host half,
 * SolidSolver.f90) and stay there.  Arrays use the Fortran memory order: (3,n) -> [3*i+k], (6,nND) -> [6*node+dof]. */
#ifndef FSILBM_SOLID_CAPI_H
#define FSILBM_SOLID_CAPI_H
#ifdef __cplusplus
extern "C" {
#endif

/* main.f90:29-48 for the structural side: read_flow_conditions + read_solid_files (inFlow.dat), allocate_solid_memory
 * (reads the mesh files named there), calculate_reference_params, set_solidbody_parameters with the root block's
 * BndConds, Initialise_solid_bodies(0).  Returns 0 or nonzero with fsolid_last_error(). */
int fsolid_open(const char *inflow_path, const int rootBC[6], int *handle);
int fsolid_close(int handle);
const char *fsolid_last_error(void);

/* out[0..15]: Lref Uref Tref Aref Fref Eref Pref nu Asfac Lchod Lspan AR denIn ntolLBM dtolLBM numsubstep */
int fsolid_flow(int handle, double out[16]);
int fsolid_nfish(int handle);
/* out[0..7]: nND nEL v_nelmts v_move iBodyModel gEQ count_Interp v_type */
int fsolid_body_info(int handle, int body, int out[8]);

int fsolid_update_pos_vel_area(int handle, int body);                                  /* UpdatePosVelArea_, Solidbody.f90:729 */
int fsolid_markers(int handle, int body, double *Exyz, double *Evel, double *Ea);       /* v_Exyz(3,n) v_Evel(3,n) v_Ea(n), out */
int fsolid_set_eforce(int handle, int body, const double *Eforce);                     /* v_Eforce(3,n), in */
int fsolid_fluid_loads(int handle, int body);                                          /* lodFlow = 0 (:911) + nodal loads (:945-967) */
int fsolid_set_lodflow(int handle, int body, const double *lodFlow);                   /* direct nodal loads (tests) */
int fsolid_structure(int handle, int body, double time, int isubstep, double deltat, double subdeltat);   /* Beam_structure, SolidSolver.f90:1820 */
int fsolid_solver(int handle, double time, int isubstep, double deltat, double subdeltat);                 /* Solver over all bodies, Solidbody.f90:386 (threads over bodies) */

/* The library's own marker arrays of a body (valid until fsolid_close; sizes 3n, 3n, n, 3n): lets a caller hand them to
 * fsilbm_ibm_interaction_force without copies. */
int fsolid_marker_ptrs(int handle, int body, double **Exyz, double **Evel, double **Ea, double **Eforce);
/* One step of host work for the listed bodies, threads over bodies: nodal loads from v_Eforce (:911,945-967), `numsubstep`
 * structural sub-steps (LBMBlockComm.f90:333-335), UpdatePosVelArea_ (:729) for the next step. */
int fsolid_advance(int handle, int nbodies, const int *bodies, double time, int numsubstep, double deltat);

/* what: 0 pos(6,nND) 1 dsp 2 vel 3 acc 4 lodFlow(gEQ) 5 lodInte(gEQ) 6 {iFish, iterNR, dnorm, cg_iterations}
 *       7 triads per element {ee(3,3) n1(3,3) n2(3,3)} row-major [i][j] = triad(i+1,j+1)   8 mss(3,nND)
 *       9 strainEnergy(2) per element after UpdateStrainEnergy   10 m_property(8) per element   11 x1(12) per element
 *       12 XYZ(3) AoA(3) UVW(3) WWW3(3)   13 lodGrav(gEQ)   14 len0,len1,geoFRM per element */
int fsolid_get(int handle, int body, int what, double *out);

/* writers (the caller's working directory holds DatBody/ DatBodySpan/ DatInfo/): what = 0 write_solid_field,
 * 1 Write_solid_v_bodies, 2 Write_solid_v_forces, 3 write_solid_Information, 4 write_information_titles,
 * 5 Write_solid_Check (appends to ./Check.dat) */
int fsolid_write(int handle, int what, double time);

#ifdef __cplusplus
}
#endif
#endif
