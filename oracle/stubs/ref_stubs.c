// TEST INFRASTRUCTURE ONLY. Link-time stand-in for the one external symbol the
// reference's sources pull in that the oracle never exercises: LAPACK's dgesdd_
// (poseutils.c:1440, procrustes only). libdogleg's entry points
// (mrcal.c:3244,6290,6435,6603,6621) are provided by the restatement in
// oracle/port/dogleg_port.c.
#include <stdio.h>

void dgesdd_(void) { fprintf(stderr, "oracle/_ref: dgesdd_() is a stub\n"); }
