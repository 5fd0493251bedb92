// SCIHarness.h -- drop-in for the header of the same name in CoGNN's Task-Worker tree (absent from /root/reference): the declarations
// the reference's engine, harness and GCN operator headers use from it, implemented on the B200 library.  See cognn_taskworker.h.
#pragma once
#include "cognn_taskworker.h"
