/* placeholder: single-state RcsGraph_* wrappers and not-yet-written batch entry points */
#include "rcsb.h"
#include "rcsb_internal.h"
extern "C" int rcsb_ik_rmr(const RcsbModel*, int, const RcsbTask*, long, double*, const double*, double, double, int, double*, double*, double*, unsigned, void*)
{ rcsb_set_error("not implemented"); return RCSB_EUNSUPPORTED; }
extern "C" bool RcsGraph_setState(RcsGraph*, const MatNd*, const MatNd*) { return false; }
