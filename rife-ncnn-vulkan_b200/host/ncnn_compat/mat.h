// ncnn_compat/mat.h -- the slice of ncnn::Mat the reference's src/main.cpp and RIFE::process touch
// (/root/reference/src/main.cpp:187,332,200-215,392-407): a ref-counted or external w x h image with
// elemsize/elempack.  Header-only, no ncnn code.
#pragma once
#include <stddef.h>
#include <stdlib.h>

#include <atomic>

namespace ncnn {

class Allocator;

class Mat {
public:
    Mat() : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0) {}
    // allocating image, main.cpp:332  `ncnn::Mat(w, h, (size_t)3, 3)`
    Mat(int _w, int _h, size_t _elemsize, int _elempack, Allocator* = 0) : Mat() { create(_w, _h, _elemsize, _elempack); }
    // external, caller-owned pixels, main.cpp:187  `ncnn::Mat(w, h, (void*)pixeldata, (size_t)3, 3)`
    Mat(int _w, int _h, void* _data, size_t _elemsize, int _elempack, Allocator* = 0)
        : data(_data), refcount(0), elemsize(_elemsize), elempack(_elempack), allocator(0), dims(2), w(_w), h(_h), d(1), c(1), cstep((size_t)_w * _h) {}
    Mat(const Mat& m)
        : data(m.data), refcount(m.refcount), elemsize(m.elemsize), elempack(m.elempack), allocator(m.allocator), dims(m.dims), w(m.w), h(m.h), d(m.d), c(m.c), cstep(m.cstep) {
        addref();
    }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m) {
        if (this == &m) return *this;
        if (m.refcount) m.refcount->fetch_add(1);
        release();
        data = m.data; refcount = m.refcount; elemsize = m.elemsize; elempack = m.elempack; allocator = m.allocator;
        dims = m.dims; w = m.w; h = m.h; d = m.d; c = m.c; cstep = m.cstep;
        return *this;
    }
    void create(int _w, int _h, size_t _elemsize, int _elempack) {
        release();
        elemsize = _elemsize; elempack = _elempack; dims = 2; w = _w; h = _h; d = 1; c = 1; cstep = (size_t)_w * _h;
        size_t bytes = (cstep * elemsize + 63) & ~(size_t)63;
        void* p = 0;
        if (bytes && posix_memalign(&p, 64, bytes + sizeof(std::atomic<int>)) == 0) {
            data = p;
            refcount = new ((unsigned char*)p + bytes) std::atomic<int>(1);
        }
    }
    void release() {
        if (refcount && refcount->fetch_sub(1) == 1) free(data);
        data = 0; refcount = 0; elemsize = 0; elempack = 0; dims = 0; w = h = d = c = 0; cstep = 0;
    }
    bool empty() const { return data == 0 || total() == 0; }
    size_t total() const { return cstep * c; }
    template <typename T> operator T*() { return (T*)data; }
    template <typename T> operator const T*() const { return (const T*)data; }

    void* data;
    std::atomic<int>* refcount;
    size_t elemsize;
    int elempack;
    Allocator* allocator;
    int dims, w, h, d, c;
    size_t cstep;

private:
    void addref() { if (refcount) refcount->fetch_add(1); }
};

}  // namespace ncnn
