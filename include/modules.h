/* modules.h -- umbrella header (empty in the reference, modules.h is 0 bytes). */
#ifndef GASR_MODULES_H
#define GASR_MODULES_H
#include "CTCBeamSearch.h"
#include "Linear.h"
#include "MemoryMonitor.h"
#include "RNN.h"
#include "RNN_Cell.h"
#include "cuMatrix.h"
#endif
