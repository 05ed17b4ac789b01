library(testthat)
library(scMergeB200)
test_check("scMergeB200")
