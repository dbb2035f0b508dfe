/* A plain-C caller of the drop-in boundary, in the style of the reference's interfaces/test/test_app.cc
 * (SolveLP: a diagonal LMI built through CONEX_NewLinearMatrixInequality / CONEX_UpdateLinearOperator).
 * Compiled as C99 against include/conex.h and linked against libconex_b200.so by tests/test_abi.py:
 * what MATLAB's loadlibrary (ConexProgram.m:28-31) and SWIG (conex.i:29) need is that the header is plain C
 * and that every symbol resolves. Exit code 0: solved (a device is present) or the library refused to create a
 * program because no device is visible; 3: anything else. */
#include <stdio.h>

#include "conex.h"

int main(void) {
  void* p = CONEX_CreateConeProgram();
  if (!p) {
    printf("no device: CONEX_CreateConeProgram returned NULL\n");
    return 0;
  }
  enum { kVars = 10 };
  int id = -1, i, status = 0;
  double b[kVars], y[kVars];
  CONEX_SolverConfiguration config;
  status |= CONEX_SetNumberOfVariables(p, kVars);
  status |= CONEX_NewLinearMatrixInequality(p, kVars, 1, &id);
  for (i = 0; i < kVars; i++) {
    status |= CONEX_UpdateLinearOperator(p, id, .3, i, i, i, 0);
    status |= CONEX_UpdateAffineTerm(p, id, .3, i, i, 0);
    b[i] = 1;
  }
  CONEX_SetDefaultOptions(&config);
  if (status != CONEX_SUCCESS) return 3;
  /* 0.3 I - 0.3 Diag(y) >= 0, maximise sum y: y = 1 */
  if (CONEX_Maximize(p, b, kVars, &config, y, kVars) != 1) return 3;
  for (i = 0; i < kVars; i++) {
    if (y[i] < 1 - 1e-3 || y[i] > 1 + 1e-6) return 3;
  }
  CONEX_DeleteConeProgram(p);
  printf("solved: y = 1\n");
  return 0;
}
