#pragma once
#include <g2o/g2o_shim.h>
