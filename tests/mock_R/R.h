/* mock of <R.h>: see Rinternals.h in this directory (test infrastructure) */
#ifndef MOCK_R_H
#define MOCK_R_H
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#endif
