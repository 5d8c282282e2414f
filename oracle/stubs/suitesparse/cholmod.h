// TEST INFRASTRUCTURE ONLY. Stand-in for <suitesparse/cholmod.h> (SuiteSparse is not in this image), so that the
// reference's uncertainty.c (which only reads the p, i, x arrays and nrow of a cholmod_sparse, uncertainty.c:866-868,966)
// compiles into oracle/_ref. The one type it needs is declared with libdogleg's stand-in.
#pragma once
#include "../dogleg.h"
