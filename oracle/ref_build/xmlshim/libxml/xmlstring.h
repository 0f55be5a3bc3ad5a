/* Empty on purpose: everything the reference needs from libxml2 is declared in tree.h
 * of this shim (oracle/ref_build/xmlshim). Test infrastructure only. */
#include "tree.h"
