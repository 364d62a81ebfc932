#include "qshim_core.h"
