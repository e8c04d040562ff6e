import sys, os, numpy as np
sys.path.insert(0,'delaydiffeq.jl_b200'); sys.path.insert(0,'tests')
import b200dde as B, oracle_lib as O
def cmp(name, model, oid, u0, p, lags, tspan, **kw):
    prob=B.DDEProblem(model,u0,None,tspan,p,constant_lags=lags)
    s=B.solve(prob,B.MethodOfSteps(B.Tsit5()),**kw)
    cfg=O.default_config(model=oid, fp_div_rule=1, **{k:v for k,v in kw.items() if k in('abstol','reltol')})
    o=O.solve_everystep(cfg,u0,p,lags,tspan[0],tspan[1])
    n=min(len(s),len(o))
    same_t=np.array_equal(s.t[:n],o.t[:n]); same_u=np.array_equal(s.u[:n],o.u[:n])
    print(name,'gpu len',len(s),'oracle len',len(o),'rc',s.retcode,o.retcode,'bit-equal t',same_t,'u',same_u, s.stats, o.stats.tolist(), 'tracked eq', s.tracked_discontinuities==o.tracked)
    if not (same_t and same_u):
        bad=[i for i in range(n) if s.t[i]!=o.t[i] or not np.array_equal(s.u[i],o.u[i])][:3]
        for i in bad: print('  first diff at',i,repr(s.t[i]),repr(o.t[i]),s.u[i],o.u[i])
cmp('logistic .3','delayed_logistic',O.MODEL_LOGISTIC,[0.1],[0.3],[1.0],(0.,50.))
cmp('logistic 1.4','delayed_logistic',O.MODEL_LOGISTIC,[0.1],[1.4],[1.0],(0.,50.))
cmp('bc','bc_model',O.MODEL_BC,[1.,1.,1.],[.2,.3,1.],[1.0],(0.,10.))
cmp('bc tight','bc_model',O.MODEL_BC,[1.,1.,1.],[.2,.3,1.],[1.0],(0.,10.),abstol=1e-10,reltol=1e-10)
cmp('mg','mackey_glass',O.MODEL_MACKEY_GLASS,[1.2],[0.2,1.2],[17.0],(0.,1000.))
cmp('1delay','constant_1delay',O.MODEL_1DELAY,[1.],[0.],[1.0],(0.,10.))
cmp('2delays','constant_2delays',O.MODEL_2DELAYS,[1.],[0.],[1/3,1/5],(0.,100.))
cmp('1delaylong','constant_1delay_long',O.MODEL_1DELAY_LONG,[1.],[0.],[0.2],(0.,100.))
cmp('backwards','backwards',O.MODEL_BACKWARDS,[1.],[0.],[-1.0],(2.,0.))
cmp('statedep','state_dependent',O.MODEL_STATE_DEP,[1.],[0.],[],(0.,10.))
cmp('1delaydep','constant_1delay_dependent',O.MODEL_1DELAY_DEP,[1.],[0.],[],(0.,10.))
