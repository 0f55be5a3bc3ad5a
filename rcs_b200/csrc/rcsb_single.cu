/* placeholder: single-state RcsGraph_* wrappers */
#include "rcsb.h"
#include "rcsb_internal.h"
extern "C" bool RcsGraph_setState(RcsGraph*, const MatNd*, const MatNd*) { return false; }
