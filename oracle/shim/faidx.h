/* oracle/shim/faidx.h -- TEST INFRASTRUCTURE ONLY: empty stand-in for the htslib header (see sam.h). */
#include "sam.h"
