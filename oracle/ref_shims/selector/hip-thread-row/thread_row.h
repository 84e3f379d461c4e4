/* Test infrastructure: stands in for the reference header of the same path (see selector_launchers.h). */
#include "selector_launchers.h"
