// stand-in (test infrastructure): forwards to the consolidated declarations
#include "nrslam_standin.h"
