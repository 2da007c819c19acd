// oracle/ref_ctx.hpp -- TEST INFRASTRUCTURE ONLY.  The context ref_driver.cpp builds from the reference's own objects, shared
// with ref_gpu_model.cpp (the drop-in shim's test driver, a separate library) and the optimizer runner of ref_optimize.hpp.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "clade.h"
#include "error_model.h"
#include "io.h"
#include "lambda.h"
#include "user_data.h"

struct ref_ctx {
    std::unique_ptr<clade> tree;
    std::unique_ptr<clade> lambda_tree;
    std::vector<const clade*> order;          // reverse level order
    int max_family_size = 0, max_root_family_size = 0;
    std::unique_ptr<error_model> em;
    user_data ud;
    input_parameters ui;
    std::string err;
};
