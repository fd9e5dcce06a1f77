/* drop-in include name of the reference (backend/radio-receiver.h) */
#include "dab_api.h"
