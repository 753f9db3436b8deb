// ncnn_compat/platform.h -- pthread wrappers with the names src/main.cpp uses (Mutex, ConditionVariable, Thread;
// /root/reference/src/main.cpp:248-292, 840-905).  Header-only.
#pragma once
#include <pthread.h>

namespace ncnn {

class Mutex {
public:
    Mutex() { pthread_mutex_init(&m_, 0); }
    ~Mutex() { pthread_mutex_destroy(&m_); }
    void lock() { pthread_mutex_lock(&m_); }
    void unlock() { pthread_mutex_unlock(&m_); }
private:
    friend class ConditionVariable;
    pthread_mutex_t m_;
};

class MutexLockGuard {
public:
    explicit MutexLockGuard(Mutex& m) : m_(m) { m_.lock(); }
    ~MutexLockGuard() { m_.unlock(); }
private:
    Mutex& m_;
};

class ConditionVariable {
public:
    ConditionVariable() { pthread_cond_init(&c_, 0); }
    ~ConditionVariable() { pthread_cond_destroy(&c_); }
    void wait(Mutex& m) { pthread_cond_wait(&c_, &m.m_); }
    void broadcast() { pthread_cond_broadcast(&c_); }
    void signal() { pthread_cond_signal(&c_); }
private:
    pthread_cond_t c_;
};

class Thread {
public:
    Thread(void* (*start)(void*), void* args = 0) { pthread_create(&t_, 0, start, args); }
    ~Thread() {}
    void join() { pthread_join(t_, 0); }
private:
    pthread_t t_;
};

}  // namespace ncnn
