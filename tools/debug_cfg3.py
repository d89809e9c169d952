import sys, os, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import chemtools_b200 as ctb
def ch4(chem, phi):
    n=len(phi); X=np.zeros((n,chem.n_species)); X[:,chem.speciesIndex("CH4")]=phi; X[:,chem.speciesIndex("O2")]=2; X[:,chem.speciesIndex("N2")]=7.52
    return chem.mass_fractions(X)
for name,rtol,atol in (("gri12",1e-9,1e-15),("gri211",1e-8,1e-14)):
    chem=ctb.ThermoKinetics(ctb.mechanism_path(name))
    rng=np.random.default_rng(20211); n=64
    T0=rng.uniform(1000,1800,n); p0=np.exp(rng.uniform(np.log(0.5),np.log(20.0),n))*101325.0; phi=rng.uniform(0.5,2.0,n)
    Y0=ch4(chem,phi)
    for linsol,clip,jac in (("inverse",1.0,"analytic"),("lu",1.0,"analytic"),("inverse",0.0,"analytic"),("inverse",1.0,"dq"),("lu",1.0,"dq")):
        b=ctb.ReactorBatch(0,chem,T0,p0,Y0,rtol,atol,1e-3,max_num_steps=100000)
        b.SetLinearSolver(linsol); b.SetJacobianClip(clip); b.SetJacobian(jac)
        code=b.Integrate(1e-3); st=b.Statistics(); status=b.Status()
        bad=np.where(status<0)[0]
        print(name,linsol,clip,jac,'code',code,'ms',round(b.LastKernelMs(),1),'nst max',st[:,0].max(),'sum',st[:,0].sum(),'ncfn',st[:,5].sum(),'netf',st[:,3].sum(),'bad',bad.tolist(), [(round(T0[i]),round(p0[i]/101325,2),round(phi[i],2)) for i in bad], st[bad].tolist(), b.Time()[bad].tolist(), flush=True)
# config1
chem=ctb.ThermoKinetics(ctb.mechanism_path("gri30"))
Y0=ch4(chem,np.array([1.0]))
for linsol,clip,jac in (("inverse",1.0,"analytic"),("lu",1.0,"analytic"),("lu",1.0,"dq"),("inverse",0.0,"analytic")):
    b=ctb.ReactorBatch(0,chem,[1200.0],[101325.0],Y0,1e-9,1e-15,0.1,max_num_steps=100000)
    b.SetLinearSolver(linsol); b.SetJacobianClip(clip); b.SetJacobian(jac)
    code=b.Integrate(0.1); st=b.Statistics()
    print('config1',linsol,clip,jac,'code',code,'ms',round(b.LastKernelMs(),1),st.tolist(),b.Time().tolist(),b.IgnitionDelay().tolist(), b.Solution()[0,0], b.Solution()[0,1:].min(), flush=True)
