// compat/lambda_lanczos.hpp — lets source written against mrcdr/lambda-lanczos compile UNCHANGED against this engine:
//     g++ -I<repo>/include -I<repo>/lambda-lanczos_b200/include -I<repo>/lambda-lanczos_b200/include/compat app.cpp -lllz
// The reference is included as <lambda_lanczos.hpp> / <lambda_lanczos/lambda_lanczos.hpp> and lives in namespace
// lambda_lanczos (lambda_lanczos.hpp:24); both spellings map onto lambda_lanczos_b200.  A host std::function mv_mul is
// accepted by the constructors (DeviceOperator<T>::host_function); pass a DeviceOperator to keep the matvec on the GPU.
#pragma once
#include "lambda_lanczos_b200/exponentiator.hpp"
#include "lambda_lanczos_b200/lambda_lanczos.hpp"
namespace lambda_lanczos = lambda_lanczos_b200;
