/* placeholder until the genozip-specific codecs are restated */
#include "oracle.h"
