/* drop-in include name of the reference (backend/radio-controller.h) */
#include "dab_api.h"
