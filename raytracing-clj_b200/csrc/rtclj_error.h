// rtclj_error.h -- the thread-local message behind rtclj_last_error(), shared by rtclj_abi.cu and
// rtclj_host.cpp so that EVERY entry point of include/rtclj_b200.h describes the error it returns.
#pragma once

// stores the formatted message for the calling thread and returns `code`
int rtclj_fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
void rtclj_error_set(const char* message);
const char* rtclj_error_get();
