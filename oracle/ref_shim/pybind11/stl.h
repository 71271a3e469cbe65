// Empty stand-in: the reference translation unit (my_cpp/common.cpp) includes this header but uses nothing from it.
