// TripleBuffer soak (shape of the reference's test/triple_buffer_test.cpp, which asserts nothing):
// a full-speed producer, a consumer that must only ever see newer frames.  CPU-only.
#include <thread>
#include <cstdio>
#include "irmv_detection/triple_buffer.hpp"
using namespace irmv_detection;
int main(){
  for (int rep=0; rep<20; rep++){
  std::array<int,3> slots={0,0,0}; TripleBuffer<int> tb(slots); std::atomic<bool> done{false};
  std::thread prod([&]{ for(int i=1;i<=200000;i++){*tb.get_producer_buffer()=i; tb.producer_commit();} while(!done.load()){*tb.get_producer_buffer()=200001; tb.producer_commit(); std::this_thread::yield();}});
  int last=0,seen=0; while(last<200000){int v=*tb.get_consumer_buffer(); if(v<last){printf("FAIL %d after %d\n",v,last); done=true; prod.join(); return 1;} last=v; seen++;}
  done=true; prod.join(); printf("ok seen %d\n",seen);} return 0; }
