#include "petscksp.h"
