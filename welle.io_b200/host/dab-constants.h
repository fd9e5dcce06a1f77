/* drop-in include name of the reference (backend/dab-constants.h) */
#include "dab_api.h"
