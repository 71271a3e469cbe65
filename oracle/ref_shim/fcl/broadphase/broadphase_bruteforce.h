#include "fcl_shim.h"
