/* test infrastructure: see ../ffstub_common.h */
#include "../ffstub_common.h"
