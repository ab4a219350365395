#pragma once
#include <cvd/image.h>
