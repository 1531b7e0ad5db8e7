#pragma once
#define checkCudaErrors(x) (x)
