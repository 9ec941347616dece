/* stand-in (TEST INFRASTRUCTURE ONLY): see RcppArmadillo.h in this directory */
#include "RcppArmadillo.h"
