#pragma once
namespace CVD { template <class T> struct Rgb { T red, green, blue; }; }
