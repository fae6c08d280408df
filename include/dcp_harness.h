/* dcp_harness.h -- C ABI of the stand-in problem builder (libdcp_harness.so, host only).
 *
 * TEST / BENCH INFRASTRUCTURE, NOT THE PRODUCT.  In a real deployment deal.II owns the mesh, the
 * DoFHandler, the AffineConstraints and the sparsity patterns (reference:
 * include/core/planet_geometry.tpp:29-120, include/core/boussinesq_model.tpp:79-412) and hands the
 * arrays below to the device library declared in dcp.h.  deal.II / Trilinos / p4est cannot be built in
 * this image, so this library restates that setup for structured meshes (cubed-sphere hypershell, unit
 * cube) and exposes every array by name.
 *
 * Spec string: comma separated key=value, keys: geometry (shell|cube), family (classic), dim (3),
 * refine, R0, R1, velocity_degree, temperature_degree, mapping_degree, patterns (0|1),
 * geometry_data (0|1), threads.
 *
 * Arrays (dtype 0=f64, 1=i32, 2=i64, 3=i8, 4=i16) -- see DESIGN.md "harness arrays" for the list:
 *   nse.l2g, temp.l2g, nse.cs.*, temp.cs.*, tab.*.phi/.dphi, q_*.w/.pts, geom.qn, geom.qt,
 *   nse.b00/.b01/.b10/.b11 .rowptr/.col, pre.b00..b11, temp.pat, nse.full, pre.full, nse.dof_xyz, ...
 */
#ifndef DCP_HARNESS_H
#define DCP_HARNESS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dcph_problem dcph_problem;

/* returns NULL on error; message via dcph_last_error() */
dcph_problem* dcph_create(const char* spec);
void dcph_destroy(dcph_problem* p);
/* 0 = ok, 1 = no such array */
int dcph_array(const dcph_problem* p, const char* name, const void** data, int64_t* count, int* dtype);
/* returns -1 if unknown */
int64_t dcph_scalar(const dcph_problem* p, const char* name);
/* number of registered arrays / name of the i-th (for enumeration) */
int dcph_n_arrays(const dcph_problem* p);
const char* dcph_array_name(const dcph_problem* p, int i);
/* same for the scalars */
int dcph_n_scalars(const dcph_problem* p);
const char* dcph_scalar_name(const dcph_problem* p, int i);
const char* dcph_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
