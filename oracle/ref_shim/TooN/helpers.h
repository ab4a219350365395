#pragma once
#include <TooN/TooN.h>
