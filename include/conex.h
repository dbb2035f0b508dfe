/* conex C ABI — the drop-in boundary of conex-b200.
 *
 * This header declares exactly the entry points, struct layouts and return conventions of the
 * reference's public C interface (reference: interfaces/conex.h:1-104, implemented in
 * interfaces/conex.cc:1-407), so that the reference's bindings (SWIG interfaces/python/conex.i,
 * MATLAB loadlibrary interfaces/matlab/util/ConexProgram.m) load libconex_b200.so unchanged.
 * Each declaration cites the reference line it replaces. Matrices are column-major FP64.
 *
 * Return conventions (preserved, including their inconsistency — reference conex.h:7-8,
 * error_checking_macros.h:15-19, cone_program.cc:532):
 *   - construction / update calls: CONEX_SUCCESS (0) or CONEX_FAILURE (1), message on stderr;
 *   - CONEX_Add* calls: the id of the new constraint (its index);
 *   - CONEX_Maximize / CONEX_Solve: 1 if solved, 0 otherwise.
 */
#ifndef CONEX_API_H
#define CONEX_API_H
#ifdef __cplusplus
extern "C" {
#endif

typedef int CONEX_STATUS;
enum { CONEX_SUCCESS = 0, CONEX_FAILURE = 1 };

/* reference conex.h:10-30 — field order and types are ABI (19 fields). */
typedef struct {
  int prepare_dual_variables;
  int initialization_mode; /* 0 cold start, 1 warm start (cone_program.h:12-15) */
  double inv_sqrt_mu_max;
  double minimum_mu;
  double maximum_mu;
  double divergence_upper_bound;
  int enable_line_search;
  double dinf_upper_bound;
  int final_centering_steps;
  double final_centering_tolerance;
  int initial_centering_steps_warmstart;
  int initial_centering_steps_coldstart;
  double warmstart_abort_threshold;
  int max_iterations;
  int iterative_refinement_iterations;
  double infeasibility_threshold;
  double kkt_error_tolerance;
  int enable_rescaling;
  int kkt_solver; /* 0 = LLT, 1 = LDLT, 2 = QR (kkt_solver.h:10-14) */
} CONEX_SolverConfiguration;

/* reference conex.h:32-35 */
typedef struct {
  double mu;
  int iteration_number;
} CONEX_IterationStats;

/* reference conex.h:37-39 */
typedef struct {
  int iterations;
} CONEX_SolutionStats;

/* reference conex.h:41-42, conex.cc:129-135 */
void* CONEX_CreateConeProgram();
void CONEX_DeleteConeProgram(void*);

/* reference conex.h:44-45, conex.cc:216-229: c - A y >= 0, A is Ar x Ac. */
int CONEX_AddDenseLinearConstraint(void* prog, const double* A, int Ar, int Ac, const double* c,
                                   int cr);

/* reference conex.h:47-49, conex.cc:190-215 (always returns -1). */
int CONEX_AddLinearInequalities(void* prog, const double* A, int Ar, int Ac, const double* lb,
                                int num_lb, const double* ub, int num_ub);

/* reference conex.h:51, conex.cc:343-355 */
int CONEX_AddQuadraticCost(void* prog, const double* A, int Ar, int Ac);

/* reference conex.h:55-57, conex.cc:137-160: Aarray holds m contiguous column-major n x n
 * matrices; Aarrayr = Aarrayc = cr = cc = n. The data is copied. Returns the constraint id. */
int CONEX_AddDenseLMIConstraint(void* prog, const double* Aarray, int Aarrayr, int Aarrayc, int m,
                                const double* cmat, int cr, int cc);

/* reference conex.h:59-61, conex.cc:162-188: as above on the 0-based variable subset `vars`. */
int CONEX_AddSparseLMIConstraint(void* prog, const double* Aarray, int Aarrayr, int Aarrayc, int m,
                                 const double* cmat, int cr, int cc, const long* vars, int vars_c);

/* reference conex.h:63-64, conex.cc:93-105: maximise b'y; writes yr = m doubles. */
int CONEX_Maximize(void* prog, const double* b, int br, const CONEX_SolverConfiguration* config,
                   double* y, int yr);

/* reference conex.h:66-67, conex.cc:107-112 */
int CONEX_Solve(void* prog, const CONEX_SolverConfiguration* config, double* y, int yr);

/* reference conex.h:69-71, conex.cc:114-127 */
void CONEX_GetDualVariable(void* prog, int i, double* x, int xr, int xc);
int CONEX_GetDualVariableSize(void* prog_ptr, int i);

/* reference conex.h:73, conex.cc:231-257 */
void CONEX_SetDefaultOptions(CONEX_SolverConfiguration* config);

/* reference conex.h:75-76, conex.cc:259-285: negative iter_num counts from the end. */
void CONEX_GetIterationStats(void* prog, CONEX_IterationStats* stats, int iter_num);

/* reference conex.h:78-80, conex.cc:365-373 */
CONEX_STATUS CONEX_UpdateLinearOperator(void* program, int constraint, double value, int variable,
                                        int row, int col, int hyper_complex_dim);

/* reference conex.h:82-84, conex.cc:287-316 */
CONEX_STATUS CONEX_NewLinearMatrixInequality(void* program, int order, int hyper_complex_dim,
                                             int* constraint_id);

/* reference conex.h:86-87, conex.cc:375-382 */
CONEX_STATUS CONEX_UpdateAffineTerm(void* program, int constraint, double value, int row, int col,
                                    int hyper_complex_dim);

/* reference conex.h:89-90, conex.cc:384-397 */
CONEX_STATUS CONEX_NewLorentzConeConstraint(void* program, int order, int* constraint_id);

/* reference conex.h:92-93, conex.cc:318-329 */
CONEX_STATUS CONEX_NewLinearInequality(void* program, int num_rows, int* constraint_id);

/* reference conex.h:95-97, conex.cc:331-363 */
CONEX_STATUS CONEX_NewQuadraticCost(void* p, int* constraint_id);
CONEX_STATUS CONEX_UpdateQuadraticCostMatrix(void* p, int id, double value, int row, int col);

/* reference conex.h:99, conex.cc:399-407 */
CONEX_STATUS CONEX_SetNumberOfVariables(void* program, int m);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif
