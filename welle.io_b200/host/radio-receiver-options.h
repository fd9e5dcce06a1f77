/* drop-in include name of the reference (backend/radio-receiver-options.h) */
#include "dab_api.h"
