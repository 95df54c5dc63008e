#include "ddp_model_wrapper.h"  // one stand-in header covers the three DDP includes of PI/mppi_controller.cuh:39-41
