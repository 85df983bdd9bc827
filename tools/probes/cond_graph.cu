// Probe: CUDA-graph WHILE conditional node driven from the device (cudaGraphSetConditional),
// the mechanism the CG local solve uses to leave its iteration loop as soon as it has converged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cond_graph cond_graph.cu && ./cond_graph
#include <cuda_runtime.h>
#include <cstdio>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void body_a(int *count) { if (threadIdx.x == 0) ++*count; }
__global__ void body_b(int *count, int limit, cudaGraphConditionalHandle h)
{
    if (threadIdx.x == 0 && *count >= limit) cudaGraphSetConditional(h, 0);
}
__global__ void prime(const int *stop, cudaGraphConditionalHandle h)
{
    if (threadIdx.x == 0) cudaGraphSetConditional(h, *stop ? 0 : 1);
}
__global__ void tail(int *count) { if (threadIdx.x == 0) *count += 1000; }

int main()
{
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int *count, *stop;
    CK(cudaMalloc(&count, 4));
    CK(cudaMalloc(&stop, 4));
    CK(cudaMemset(count, 0, 4));
    CK(cudaMemset(stop, 0, 4));
    cudaGraph_t g;
    CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    // head: captured into the main graph
    CK(cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    prime<<<1, 32, 0, st>>>(stop, h);
    // the conditional node, added to the capture's current dependency set
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t *deps;
    size_t ndeps;
    cudaGraph_t gcap;
    CK(cudaStreamGetCaptureInfo(st, &status, nullptr, &gcap, &deps, &ndeps));
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, gcap, deps, ndeps, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    CK(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
    tail<<<1, 32, 0, st>>>(count);
    CK(cudaStreamEndCapture(st, &g));
    // body: its own capture
    cudaStream_t st2;
    CK(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
    CK(cudaStreamBeginCaptureToGraph(st2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    body_a<<<1, 32, 0, st2>>>(count);
    body_b<<<1, 32, 0, st2>>>(count, 37, h);
    CK(cudaStreamEndCapture(st2, nullptr));
    cudaGraphExec_t ex;
    CK(cudaGraphInstantiate(&ex, g, 0));
    CK(cudaGraphLaunch(ex, st));
    CK(cudaStreamSynchronize(st));
    int hcount = -1;
    CK(cudaMemcpy(&hcount, count, 4, cudaMemcpyDeviceToHost));
    printf("count after while-graph = %d (expect 1037)\n", hcount);
    // second launch with stop preset: zero trips
    int one = 1;
    CK(cudaMemcpy(stop, &one, 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(count, 0, 4));
    CK(cudaGraphLaunch(ex, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(&hcount, count, 4, cudaMemcpyDeviceToHost));
    printf("count with stop preset = %d (expect 1000)\n", hcount);
    // timing: 1000 trips of two tiny kernels
    CK(cudaMemset(stop, 0, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%s\n", hcount == 1000 ? "COND_GRAPH_OK" : "COND_GRAPH_BAD");
    return 0;
}
