// compat/exponentiator.hpp — see compat/lambda_lanczos.hpp (the reference's <exponentiator.hpp>).
#pragma once
#include "lambda_lanczos.hpp"
