/*
 * amg_setup_b200.h -- the reference's own setup interface, served by the B200 engine.
 *
 * libamg_setup_b200.so exports, with the reference's names, signatures and types,
 *
 *     void amg_setup (uint n, const uint *Ai, const uint *Aj, const double *Av,
 *                     struct amg_setup_data *data);             amg_setup.h:5, amg_setup.c:60
 *     void amg_export(struct amg_setup_data *data);             amg_setup.h:9, amg_setup.c:405
 *     void free_data (struct amg_setup_data **data);            amg_setup.h,   amg_setup.c:3487
 *
 * so that the reference's driver (serial_amg.c:98-104) -- or any caller of those three --
 * links against it UNMODIFIED instead of against amg_setup.o/amg_tools.o.  amg_setup runs
 * the hierarchy construction on the GPU (libomp_amg_b200.so, include/omp_amg_b200.h) and
 * fills a real struct amg_setup_data on the host (every field of amg_tools.h:29-52, arrays
 * from malloc so that the caller may free them as the reference's free_data does);
 * amg_export writes amg.dat, amg_W.dat, amg_AfP.dat and amg_Aff.dat into the current
 * directory, byte for byte what the reference writes.
 *
 * Integer width.  gslib fixes "uint" at build time (types.h:52-64).  The reference's Makefile
 * builds with -DUSE_LONG -DGLOBAL_LONG, i.e. uint = unsigned long, and that is how the shim in
 * this repo is built (omp_amg_b200/csrc/Makefile: SHIM_DEFS).  A gslib built without those
 * flags needs the shim rebuilt with the same flags; the types below follow the same macros.
 *
 * A caller that includes the reference's headers does not need this file: the structs below
 * are the mirror of amg_tools.h used to compile the shim without the reference tree.
 */
#ifndef AMG_SETUP_B200_H
#define AMG_SETUP_B200_H

#if defined(USE_LONG_LONG)
typedef unsigned long long amgb_uint;
#elif defined(USE_LONG)
typedef unsigned long amgb_uint;
#else
typedef unsigned int amgb_uint;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* struct csr_mat, amg_tools.h:5 */
struct amgb_csr_mat {
  amgb_uint rn, cn, *row_off, *col;
  double *a;
};

/* struct amg_setup_data, amg_tools.h:29-52 (same field order and types) */
struct amgb_amg_setup_data {
  double tolc;
  double gamma;
  double *n;
  double *nnz;
  double *nnzf;
  double *nnzfp;
  double *m;
  double *rho;
  struct amgb_csr_mat **A;
  amgb_uint *id;
  amgb_uint **idc;
  amgb_uint **idf;
  double **C;
  double **F;
  double **D;
  struct amgb_csr_mat **Af;
  struct amgb_csr_mat **W;
  struct amgb_csr_mat **AfP;
  amgb_uint nlevels;
  amgb_uint nullspace;
};

#ifndef AMG_SETUP            /* the reference's amg_setup.h already declares these */
struct amg_setup_data;
void amg_setup(amgb_uint n, const amgb_uint *Ai, const amgb_uint *Aj, const double *Av,
               struct amg_setup_data *data);
void amg_export(struct amg_setup_data *data);
void free_data(struct amg_setup_data **data);
#endif

#ifdef __cplusplus
}
#endif
#endif
