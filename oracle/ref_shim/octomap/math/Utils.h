#include "../octomap.h"
