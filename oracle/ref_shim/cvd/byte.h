#pragma once
namespace CVD { typedef unsigned char byte; }
