/* Test infrastructure: the selector (adaptive.cpp) includes the HIP runtime but uses nothing of it. */
