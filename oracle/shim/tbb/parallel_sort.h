#include "tbb.h"
